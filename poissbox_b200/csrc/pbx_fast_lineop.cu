// pbx_fast_lineop.cu -- FAST formulation of ONE 1-D compact operator on every line of a brick:
//   grad_1d / div_1d / interp_1d / interp_1d_div  (src/compact_schemes.f90:155-329), i.e.
//   P = A^-1 B  with B the 4-point right-hand side of eval_1d_rhs (:332-372) and
//   A = circ[al, 1, al] = (1 - r E^-1)(1 - r E)/(1 + r^2).
// Same machinery as the Laplacian passes (pbx_fast_common.cuh): a thread owns a 16-point chunk of
// a line in registers, applies the stencil (in the reference's own form a (f_i +- f_i-1) +
// b (f_i+1 +- f_i-2), so constants differentiate to exactly zero), runs the causal first-order
// recursion from zero state, corrects it with the look-back over the preceding chunks' end states,
// and the same anti-causally; (1 + r^2) is folded into a and b.  16 B/point of HBM traffic per
// operator, against three global-memory sweeps for the REFERENCE schedule's thread-per-line Thomas.
// Used by the FAST schedule of grad / div / interp (8 + 8 + 3 line operators, reference stage
// order); results agree with the REFERENCE schedule to rounding (tests/test_parity_gpu.py).
#include "pbx_fast_common.cuh"

namespace pbx {

using namespace fast;

namespace {

constexpr int XW = 8;
constexpr int CPAD = LC + 2;

struct LineOp {
    CompositeCoef cc;     // only r, pw, look, nlook are used
    double a, b;          // right-hand-side coefficients times (1 + r^2)
    int deriv;            // 1: opsign -1 (differences), 0: opsign +1 (sums)
    int shift;            // 0: stagger -1 (cell -> vertex), 1: stagger +1 (vertex -> cell)
};

// rhs_k = a (f_{k+sh} +- f_{k-1+sh}) + b (f_{k+1+sh} +- f_{k-2+sh}),  e[k+3] = f_k
__device__ __forceinline__ void stencil4(const LineOp &op, const double (&e)[LC + 6], double (&o)[LC])
{
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        const double f0 = op.shift ? e[k + 4] : e[k + 3], f1 = op.shift ? e[k + 3] : e[k + 2];
        const double f2 = op.shift ? e[k + 5] : e[k + 4], f3 = op.shift ? e[k + 2] : e[k + 1];
        const double t1 = op.deriv ? f0 - f1 : f0 + f1;
        const double t2 = op.deriv ? f2 - f3 : f2 + f3;
        o[k] = fma(op.b, t2, op.a * t1);
    }
}

// single-pole look-back: S = sum_m r^(16 (m-1)) E_(t -+ m)
__device__ __forceinline__ double lookback1(const CompositeCoef &c, const Xchg &x, int slot, int dir)
{
    double S = x.get(slot, x.nb(dir));
#pragma unroll
    for (int m = 2; m <= MAXLOOK; ++m)
        if (m <= c.nlook) S = fma(c.look[m - 1], x.get(slot, x.nb(dir * m)), S);
    return S;
}

// v <- A^-1 v (up to the folded factor); slots s0, s0+1; two barriers
template <class Bar>
__device__ __forceinline__ void solve1_chunk(const CompositeCoef &c, const Xchg &x, int s0,
                                             double (&v)[LC], Bar bar)
{
    double y = 0.0;
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        y = fma(c.r, y, v[k]);
        v[k] = y;
    }
    x.put(s0, y);
    bar();
    const double S = lookback1(c, x, s0, -1);
    double w = 0.0;
#pragma unroll
    for (int k = LC - 1; k >= 0; --k) {
        const double yk = fma(c.pw[k], S, v[k]);   // corrected causal value
        w = fma(c.r, w, yk);
        v[k] = w;
    }
    x.put(s0 + 1, w);
    bar();
    const double W = lookback1(c, x, s0 + 1, +1);
#pragma unroll
    for (int k = 0; k < LC; ++k) v[k] = fma(c.pw[LC - 1 - k], W, v[k]);
}

struct LYZ {
    LineOp op;
    int nx, T, ng;
    long long sl, sg;
    SegGeom seg;          // lines of more than 512 points: blockIdx.z numbers the segment
};

__global__ void __launch_bounds__(NT, 3)
lineop_yz_kernel(const __grid_constant__ LYZ p, const double *__restrict__ in, double *__restrict__ out)
{
    __shared__ double sm[8 * NT];
    const int tx = threadIdx.x, t = threadIdx.y, tz = threadIdx.z;
    const int x = blockIdx.x * XW + tx;
    const int g = blockIdx.y * blockDim.z + tz;
    const bool live = (x < p.nx) && (g < p.ng);
    const SegChunk sc = seg_chunk(p.seg, blockIdx.z, t);
    Xchg xc{sm, (tz * p.T + t) * XW + tx, t, p.T, XW, p.seg.nseg > 1 ? 1 : 0};
    const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;
    double f[LC];
#pragma unroll
    for (int k = 0; k < LC; ++k) f[k] = live ? __ldg(in + base + k * p.sl) : 0.0;
    put_halo(xc, 2, f);
    __syncthreads();
    double e[LC + 6], v[LC];
    get_halo(xc, 2, f, e);
    stencil4(p.op, e, v);
    solve1_chunk(p.op.cc, xc, 0, v, BarAll());
    if (live && sc.interior) {
#pragma unroll
        for (int k = 0; k < LC; ++k) out[base + k * p.sl] = v[k];
    }
}

struct LX {
    LineOp op;
    int n, T, R;
    long long nlines;
};

__global__ void __launch_bounds__(NT, 2)
lineop_x_kernel(const __grid_constant__ LX p, const double *__restrict__ in, double *__restrict__ out)
{
    extern __shared__ double sm[];
    const int t = threadIdx.x, r = threadIdx.y;
    const int tid = r * p.T + t, nthr = p.T * p.R;
    const long long line0 = (long long)blockIdx.x * p.R;
    const int rlive = (int)((p.nlines - line0) < p.R ? (p.nlines - line0) : p.R);
    const int rowpad = p.T * CPAD;
    double *buf = sm + 8 * NT;
    Xchg xc{sm, tid, t, p.T, 1};
    {
        const double2 *src = reinterpret_cast<const double2 *>(in + line0 * p.n);
        const int n2 = rlive * p.n / 2;
        for (int q = tid; q < n2; q += nthr) {
            double2 val = __ldg(src + q);
            int i = 2 * q, rr = i / p.n, ii = i - rr * p.n;
            *reinterpret_cast<double2 *>(buf + rr * rowpad + (ii >> 4) * CPAD + (ii & 15)) = val;
        }
    }
    __syncthreads();
    double f[LC];
    {
        const double2 *src = reinterpret_cast<const double2 *>(buf + r * rowpad + t * CPAD);
#pragma unroll
        for (int k = 0; k < LC / 2; ++k) {
            double2 val = (r < rlive) ? src[k] : make_double2(0.0, 0.0);
            f[2 * k] = val.x;
            f[2 * k + 1] = val.y;
        }
    }
    put_halo(xc, 2, f);
    __syncthreads();
    double e[LC + 6], v[LC];
    get_halo(xc, 2, f, e);
    stencil4(p.op, e, v);
    solve1_chunk(p.op.cc, xc, 0, v, BarAll());
    {
        double2 *d0 = reinterpret_cast<double2 *>(buf + r * rowpad + t * CPAD);
#pragma unroll
        for (int k = 0; k < LC / 2; ++k) d0[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
    __syncthreads();
    {
        double2 *dst = reinterpret_cast<double2 *>(out + line0 * p.n);
        const int n2 = rlive * p.n / 2;
        for (int q = tid; q < n2; q += nthr) {
            int i = 2 * q, rr = i / p.n, ii = i - rr * p.n;
            dst[q] = *reinterpret_cast<const double2 *>(buf + rr * rowpad + (ii >> 4) * CPAD + (ii & 15));
        }
    }
}

}  // namespace

// one 1-D compact operator along dir on a brick the FAST schedule supports (fast_supported())
int fast_line_op(cudaStream_t s, const Brick &g, int dir, OpKind kind, int stagger, double dx,
                 const double *in, double *out, long long *launches)
{
    if (in == out || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)) {
        set_last_error("FAST line operators need distinct, 16-byte aligned input and output");
        return PBX_ERR_ARG;
    }
    LineOp op;
    make_composite_coef(kind, dx, &op.cc);
    double a, b;
    scheme_ab(kind, dx, &a, &b);
    const double sc = 1.0 + op.cc.r * op.cc.r;
    op.a = a * sc;
    op.b = b * sc;
    op.deriv = kind == OP_DERIV;
    op.shift = stagger == PBX_STAGGER_BACKWARD ? 0 : 1;
    if (dir == 0) {
        LX p;
        p.op = op;
        p.n = g.nx;
        p.T = g.nx / LC;
        p.R = NT / p.T;
        p.nlines = (long long)g.ny * g.nz;
        if (p.R > p.nlines) p.R = (int)p.nlines;
        const size_t smem = sizeof(double) * (8 * NT + (size_t)p.R * p.T * CPAD);
        static bool attr_set[64] = {false};
        int dev_ = 0;
        cudaGetDevice(&dev_);
        if (!attr_set[dev_ & 63]) {
            PBX_CUDA(cudaFuncSetAttribute(lineop_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          64 * 1024));
            attr_set[dev_ & 63] = true;
        }
        lineop_x_kernel<<<(unsigned)((p.nlines + p.R - 1) / p.R), dim3(p.T, p.R), smem, s>>>(p, in, out);
    } else {
        LYZ p;
        p.op = op;
        const int n = dir == 1 ? g.ny : g.nz;
        p.nx = g.nx;
        p.seg = seg_geometry(n / LC);
        p.T = p.seg.T;
        p.ng = dir == 1 ? g.nz : g.ny;
        p.sl = dir == 1 ? (long long)g.nx : (long long)g.nx * g.ny;
        p.sg = dir == 1 ? (long long)g.nx * g.ny : (long long)g.nx;
        int G = NT / (XW * p.T);
        if (G < 1) G = 1;
        if (G > p.ng) G = p.ng;
        dim3 block(XW, p.T, G), grid((g.nx + XW - 1) / XW, (p.ng + G - 1) / G, p.seg.nseg);
        lineop_yz_kernel<<<grid, block, 0, s>>>(p, in, out);
    }
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
