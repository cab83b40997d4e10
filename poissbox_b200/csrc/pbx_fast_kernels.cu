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
#include "pbx_internal.h"

namespace pbx {

namespace {

constexpr int XW = 8;             // x-width of a y/z-pass tile (64-byte row segments)
constexpr int NT = 256;           // threads per CTA
constexpr int CPAD = LC + 2;      // x pass: doubles per chunk in the padded staging buffer

// ---------------------------------------------------------------------------------------------
// register-array building blocks (everything unrolled, indices compile-time)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stencil1(const CompositeCoef &c, const double (&e)[LC + 6],
                                         double (&o)[LC])
{
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        double s1 = e[k + 2] + e[k + 4];
        double s2 = e[k + 1] + e[k + 5];
        double s3 = e[k] + e[k + 6];
        o[k] = fma(c.c3, s3, fma(c.c2, s2, fma(c.c1, s1, c.c0 * e[k + 3])));
    }
}

// two stencils of the same input share the symmetric sums
__device__ __forceinline__ void stencil2(const CompositeCoef &ca, const CompositeCoef &cb,
                                         const double (&e)[LC + 6], double (&oa)[LC],
                                         double (&ob)[LC])
{
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        double s1 = e[k + 2] + e[k + 4];
        double s2 = e[k + 1] + e[k + 5];
        double s3 = e[k] + e[k + 6];
        oa[k] = fma(ca.c3, s3, fma(ca.c2, s2, fma(ca.c1, s1, ca.c0 * e[k + 3])));
        ob[k] = fma(cb.c3, s3, fma(cb.c2, s2, fma(cb.c1, s1, cb.c0 * e[k + 3])));
    }
}

__device__ __forceinline__ void fwd_local(double r, double (&v)[LC], double &ey, double &ez)
{
    double y = 0.0, z = 0.0;
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        y = fma(r, y, v[k]);
        z = fma(r, z, y);
        v[k] = z;
    }
    ey = y;
    ez = z;
}

__device__ __forceinline__ void bwd_local(double r, double (&v)[LC], double &ew, double &ex)
{
    double w = 0.0, x = 0.0;
#pragma unroll
    for (int k = LC - 1; k >= 0; --k) {
        w = fma(r, w, v[k]);
        x = fma(r, x, w);
        v[k] = x;
    }
    ew = w;
    ex = x;
}

// homogeneous correction for a true incoming causal state (Y, Z) = (y_-1, z_-1)
__device__ __forceinline__ void fwd_fix(const CompositeCoef &c, double (&v)[LC], double Y, double Z)
{
#pragma unroll
    for (int k = 0; k < LC; ++k) v[k] = fma(c.pw[k], fma((double)(k + 1), Y, Z), v[k]);
}

// ... and for a true incoming anti-causal state (W, X) = (w_LC, x_LC)
__device__ __forceinline__ void bwd_fix(const CompositeCoef &c, double (&v)[LC], double W, double X)
{
#pragma unroll
    for (int k = 0; k < LC; ++k)
        v[k] = fma(c.pw[LC - 1 - k], fma((double)(LC - k), W, X), v[k]);
}

// Shared-memory exchange area: slot s of thread q lives at sm[s*NT + q]; the chunk t' of the same
// line belongs to thread q + (t' - t)*tstride.
struct Xchg {
    double *sm;
    int q, t, T, tstride;
    __device__ __forceinline__ int nb(int dt) const
    {
        int tt = t + dt;
        tt %= T;
        if (tt < 0) tt += T;
        return q + (tt - t) * tstride;
    }
    __device__ __forceinline__ void put(int slot, double v) const { sm[slot * NT + q] = v; }
    __device__ __forceinline__ double get(int slot, int qq) const { return sm[slot * NT + qq]; }
};

// true incoming state from the local end states published in slots (sy, sz); dir = -1 looks at
// chunks t-1, t-2, ... (causal), dir = +1 at t+1, t+2, ... (anti-causal)
__device__ __forceinline__ void lookback(const CompositeCoef &c, const Xchg &x, int sy, int sz,
                                         int dir, double &Y, double &Z)
{
    int q1 = x.nb(dir);
    Y = x.get(sy, q1);
    Z = x.get(sz, q1);
#pragma unroll
    for (int m = 2; m <= MAXLOOK; ++m) {
        if (m <= c.nlook) {
            int qm = x.nb(dir * m);
            double ey = x.get(sy, qm), ez = x.get(sz, qm);
            double p = c.look[m - 1];
            Y = fma(p, ey, Y);
            Z = fma(p, fma((double)(LC * (m - 1)), ey, ez), Z);
        }
    }
}

// publish the first and last three points of a chunk (slots s0 .. s0+5)
__device__ __forceinline__ void put_halo(const Xchg &x, int s0, const double (&v)[LC])
{
    x.put(s0 + 0, v[0]);
    x.put(s0 + 1, v[1]);
    x.put(s0 + 2, v[2]);
    x.put(s0 + 3, v[LC - 3]);
    x.put(s0 + 4, v[LC - 2]);
    x.put(s0 + 5, v[LC - 1]);
}

// assemble e = { last 3 of chunk t-1, v, first 3 of chunk t+1 }
__device__ __forceinline__ void get_halo(const Xchg &x, int s0, const double (&v)[LC],
                                         double (&e)[LC + 6])
{
    int ql = x.nb(-1), qr = x.nb(+1);
    e[0] = x.get(s0 + 3, ql);
    e[1] = x.get(s0 + 4, ql);
    e[2] = x.get(s0 + 5, ql);
#pragma unroll
    for (int k = 0; k < LC; ++k) e[k + 3] = v[k];
    e[LC + 3] = x.get(s0 + 0, qr);
    e[LC + 4] = x.get(s0 + 1, qr);
    e[LC + 5] = x.get(s0 + 2, qr);
}

// solve NF right-hand sides in registers: v[f] <- A_f^-2 v[f].  Uses slots [s0, s0 + 4*NF); two
// __syncthreads.  Every thread of the CTA must call it.
template <int NF>
__device__ __forceinline__ void solve_chunks(const CompositeCoef *const (&c)[NF], const Xchg &x,
                                             int s0, double (&v)[NF][LC])
{
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double ey, ez;
        fwd_local(c[f]->r, v[f], ey, ez);
        x.put(s0 + 2 * f, ey);
        x.put(s0 + 2 * f + 1, ez);
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double Y, Z, ew, ex;
        lookback(*c[f], x, s0 + 2 * f, s0 + 2 * f + 1, -1, Y, Z);
        fwd_fix(*c[f], v[f], Y, Z);
        bwd_local(c[f]->r, v[f], ew, ex);
        x.put(s0 + 2 * NF + 2 * f, ew);
        x.put(s0 + 2 * NF + 2 * f + 1, ex);
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double W, X;
        lookback(*c[f], x, s0 + 2 * NF + 2 * f, s0 + 2 * NF + 2 * f + 1, +1, W, X);
        bwd_fix(*c[f], v[f], W, X);
    }
}

// ---------------------------------------------------------------------------------------------
// y / z pass.  Tile: XW consecutive x, T chunks (one whole line), G lines in the remaining
// direction.  blockDim = (XW, T, G).  sl = stride along the line, sg = stride between lines.
// ---------------------------------------------------------------------------------------------
struct YZParams {
    CompositeCoef M, D;
    int nx, T, ng;          // x extent, chunks per line, number of lines in the "group" direction
    long long sl, sg;       // strides (in doubles) along the line / between lines of a group
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
    Xchg xc{sm, (tz * p.T + t) * XW + tx, t, p.T, XW};
    const long long base = (long long)x + (long long)(t * LC) * p.sl + (long long)g * p.sg;

    double a[LC], b[LC];
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        a[k] = live ? __ldg(A + base + k * p.sl) : 0.0;
        b[k] = live ? __ldg(B + base + k * p.sl) : 0.0;
    }
    put_halo(xc, 0, a);
    put_halo(xc, 6, b);
    __syncthreads();

    double v[3][LC];   // v0 = S_M a, v1 = S_D b, v2 = S_M b
    {
        double e[LC + 6];
        get_halo(xc, 0, a, e);
        stencil1(p.M, e, v[0]);
        get_halo(xc, 6, b, e);
        stencil2(p.D, p.M, e, v[1], v[2]);
    }
    const CompositeCoef *const cs[3] = {&p.M, &p.D, &p.M};
    solve_chunks<3>(cs, xc, 12, v);

    if (live) {
#pragma unroll
        for (int k = 0; k < LC; ++k) {
            C[base + k * p.sl] = v[0][k] + v[1][k];
            Dd[base + k * p.sl] = v[2][k];
        }
    }
}

// z pass:  out = M c + D d ; optionally partial sums of pv . out (one per CTA, deterministic)
__global__ void __launch_bounds__(NT, 2)
zpass_kernel(const __grid_constant__ YZParams p, const double *__restrict__ Cc,
             const double *__restrict__ Dd, double *__restrict__ out,
             const double *__restrict__ pv, double *__restrict__ partials)
{
    extern __shared__ double sm[];
    const int tx = threadIdx.x, t = threadIdx.y, tz = threadIdx.z;
    const int x = blockIdx.x * XW + tx;
    const int g = blockIdx.y * blockDim.z + tz;
    const bool live = (x < p.nx) && (g < p.ng);
    Xchg xc{sm, (tz * p.T + t) * XW + tx, t, p.T, XW};
    const long long base = (long long)x + (long long)(t * LC) * p.sl + (long long)g * p.sg;

    double v[2][LC];
    {
        double c[LC], d[LC];
#pragma unroll
        for (int k = 0; k < LC; ++k) {
            c[k] = live ? __ldg(Cc + base + k * p.sl) : 0.0;
            d[k] = live ? __ldg(Dd + base + k * p.sl) : 0.0;
        }
        put_halo(xc, 0, c);
        put_halo(xc, 6, d);
        __syncthreads();
        double e[LC + 6];
        get_halo(xc, 0, c, e);
        stencil1(p.M, e, v[0]);
        get_halo(xc, 6, d, e);
        stencil1(p.D, e, v[1]);
    }
    const CompositeCoef *const cs[2] = {&p.M, &p.D};
    solve_chunks<2>(cs, xc, 12, v);

    double dot = 0.0;
    if (live) {
        if (pv != nullptr) {
#pragma unroll
            for (int k = 0; k < LC; ++k) {
                double o = v[0][k] + v[1][k];
                dot = fma(__ldg(pv + base + k * p.sl), o, dot);
                out[base + k * p.sl] = o;
            }
        } else {
#pragma unroll
            for (int k = 0; k < LC; ++k) out[base + k * p.sl] = v[0][k] + v[1][k];
        }
    }
    if (pv != nullptr) {
        // fixed-shape (deterministic) block reduction through the exchange area
        const int lin = xc.q;
        const int nthr = blockDim.x * blockDim.y * blockDim.z;
        __syncthreads();   // exchange area is free again
        sm[lin] = dot;
        __syncthreads();
        double s = 0.0;
        if (lin < 32)
            for (int i = lin; i < nthr; i += 32) s += sm[i];
        __syncthreads();
        if (lin < 32) sm[lin] = s;
        __syncthreads();
        if (lin == 0) {
            double tot = 0.0;
            const int m = nthr < 32 ? nthr : 32;
            for (int i = 0; i < m; ++i) tot += sm[i];
            partials[blockIdx.y * gridDim.x + blockIdx.x] = tot;
        }
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
    double *xs = sm;                         // exchange area: 14 slots
    double *buf0 = sm + 14 * NT;             // staging, R * rowpad doubles
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
    put_halo(xc, 0, f);
    __syncthreads();
    double v[2][LC];   // v0 = S_D f , v1 = S_M f
    {
        double e[LC + 6];
        get_halo(xc, 0, f, e);
        stencil2(p.D, p.M, e, v[0], v[1]);
    }
    const CompositeCoef *const cs[2] = {&p.D, &p.M};
    solve_chunks<2>(cs, xc, 6, v);

    // stage out through the padded buffers, then coalesced 16-byte stores
    {
        double2 *d0 = reinterpret_cast<double2 *>(buf0 + r * rowpad + t * CPAD);
        double2 *d1 = reinterpret_cast<double2 *>(buf1 + r * rowpad + t * CPAD);
#pragma unroll
        for (int k = 0; k < LC / 2; ++k) {
            d0[k] = make_double2(v[0][2 * k], v[0][2 * k + 1]);
            d1[k] = make_double2(v[1][2 * k], v[1][2 * k + 1]);
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

// host helpers ---------------------------------------------------------------------------------
inline int pow2_floor(int v)
{
    int p = 1;
    while (2 * p <= v) p *= 2;
    return p;
}

constexpr int YZ_SLOTS = 12 + 12;   // 2 halos + up to 3 ops x (2 fwd + 2 bwd)
constexpr size_t YZ_SMEM = sizeof(double) * YZ_SLOTS * NT;

}  // namespace

bool fast_supported(int nx, int ny, int nz)
{
    // chunks of 16 points; one CTA must hold a whole line of chunks (T <= 64 at XW = 8 for y/z,
    // T <= 256 for x)
    auto ok = [](int n, int tmax) { return n >= LC && n % LC == 0 && n / LC <= tmax; };
    return ok(nx, NT) && ok(ny, NT / XW) && ok(nz, NT / XW);
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
    const size_t smem = sizeof(double) * (14 * NT + 2 * (size_t)p.R * p.T * CPAD);
    static bool attr_set = false;
    if (!attr_set) {
        PBX_CUDA(cudaFuncSetAttribute(xpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      110 * 1024));
        attr_set = true;
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
    p->T = n / LC;
    p->ng = dir == 1 ? g.nz : g.ny;
    p->sl = dir == 1 ? (long long)g.nx : (long long)g.nx * g.ny;
    p->sg = dir == 1 ? (long long)g.nx * g.ny : (long long)g.nx;
    int G = NT / (XW * p->T);
    if (G < 1) G = 1;
    if (G > p->ng) G = p->ng;
    *block = dim3(XW, p->T, G);
    *grid = dim3((g.nx + XW - 1) / XW, (p->ng + G - 1) / G);
}

int fast_ypass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *A,
               const double *B, double *C, double *D, long long *launches)
{
    YZParams p;
    p.M = fc.M;
    p.D = fc.D[1];
    dim3 grid, block;
    yz_geometry(g, 1, &p, &grid, &block);
    static bool attr_set = false;
    if (!attr_set) {
        PBX_CUDA(cudaFuncSetAttribute(ypass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)YZ_SMEM));
        attr_set = true;
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
    return (int)(grid.x * grid.y);
}

int fast_zpass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *C,
               const double *D, double *out, const double *pvec, double *dot_partials,
               int *n_partials, long long *launches)
{
    YZParams p;
    p.M = fc.M;
    p.D = fc.D[2];
    dim3 grid, block;
    yz_geometry(g, 2, &p, &grid, &block);
    static bool attr_set = false;
    if (!attr_set) {
        PBX_CUDA(cudaFuncSetAttribute(zpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)YZ_SMEM));
        attr_set = true;
    }
    zpass_kernel<<<grid, block, YZ_SMEM, s>>>(p, C, D, out, pvec, dot_partials);
    if (n_partials) *n_partials = (int)(grid.x * grid.y);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
