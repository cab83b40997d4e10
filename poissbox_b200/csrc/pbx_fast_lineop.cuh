// pbx_fast_lineop.cuh -- the arithmetic of ONE 1-D compact operator P = A^-1 B on a 16-point chunk in
// registers (see pbx_fast_lineop.cu), shared by the generic line-operator kernels and their
// TMA-pipelined variants (pbx_fast_tma.cu): same operations, same bits.
#pragma once

#include "pbx_fast_common.cuh"

namespace pbx {
namespace lineop {

using namespace fast;

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

LineOp make_line_op(OpKind kind, int stagger, double dx);

}  // namespace lineop
}  // namespace pbx
