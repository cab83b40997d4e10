// pbx_fast_common.cuh -- register-array building blocks of the FAST schedule, shared by the
// TMA-pipelined kernels (pbx_fast_tma.cu) and the generic fallback kernels (pbx_fast_kernels.cu).
//
// A composite 1-D operator O = A^-2 S (pbx_internal.h) acts on a periodic line cut into chunks of
// LC = 16 points; a thread owns one chunk of one line in registers.
//
//   derivative composite  (D = D+ G-):  STENCIL FIRST.  S_D is evaluated in difference form
//        S_D f = c1 (s1 - 2 f0) + c2 (s2 - 2 f0) + c3 (s3 - 2 f0),  s_k = f_k + f_-k,
//     so that a constant gives exactly zero, as the reference's f(i) - f(i-1) does
//     (tests/lapl/test_lapl.f90:57-74); then the double recursion.
//   interpolation composite (M = I+ I-):  SOLVE FIRST, then S_M on the solved values.  A^-2
//     amplifies grid-scale noise by up to (1 - 2 al)^-2 = 6.25 while S_M annihilates it, so doing
//     the smoothing stencil last keeps the rounding noise that the derivative operators of the
//     other axes later amplify by ~4/dx^2 about 6x smaller (measured against a long-double
//     evaluation: same error as the reference's own order of operations).
//
// The recursion: local causal sweep from zero state, y_k = s_k + r y_{k-1}, z_k = y_k + r z_{k-1};
// the true incoming state is assembled from the neighbours' local end states (look-back over
// nlook chunks, the influence of chunk t-m decaying as r^(16 m)); the chunk is corrected with the
// homogeneous solution r^(k+1) (Z + (k+1) Y); then the same anti-causally.  Periodicity needs no
// Sherman-Morrison step: the look-back wraps around the line.
#pragma once

#include <atomic>

#include "pbx_internal.h"
#include "pbx_ptx.cuh"

namespace pbx {
namespace fast {

constexpr int NT = 256;   // compute threads per CTA

template <bool DIFF>
__device__ __forceinline__ double stencil_point(const CompositeCoef &c, double f0, double s1,
                                                double s2, double s3)
{
    if (DIFF) {
        double d1 = fma(-2.0, f0, s1), d2 = fma(-2.0, f0, s2), d3 = fma(-2.0, f0, s3);
        return fma(c.c3, d3, fma(c.c2, d2, c.c1 * d1));
    }
    return fma(c.c3, s3, fma(c.c2, s2, fma(c.c1, s1, c.c0 * f0)));
}

// o = S e, e = { 3 halo points, LC chunk points, 3 halo points }
template <bool DIFF>
__device__ __forceinline__ void stencil(const CompositeCoef &c, const double (&e)[LC + 6],
                                        double (&o)[LC])
{
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        double s1 = e[k + 2] + e[k + 4];
        double s2 = e[k + 1] + e[k + 5];
        double s3 = e[k] + e[k + 6];
        o[k] = stencil_point<DIFF>(c, e[k + 3], s1, s2, s3);
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

// Shared-memory exchange area: slot s of thread q lives at sm[s*NT + q]; chunk t' of the same
// line belongs to thread q + (t' - t)*tstride.
// On an OPEN line (one slab of a z-decomposed box) chunks outside [0, T) do not exist: nb() returns
// -1 for them and get() reads zero (zero recursion state, zero stencil halo); the coupling to the
// neighbouring slabs is added afterwards as a low-rank correction (pbx_dist_tables.cu).
struct Xchg {
    double *sm;
    int q, t, T, tstride;
    int open = 0;
    int dead = 0;   // a thread without a chunk (tiles whose lines do not fill the CTA): reads zeros
    __device__ __forceinline__ int nb(int dt) const
    {
        if (dead) return -1;
        int tt = t + dt;
        if (open) {
            if (tt < 0 || tt >= T) return -1;
        } else {
            tt %= T;
            if (tt < 0) tt += T;
        }
        return q + (tt - t) * tstride;
    }
    __device__ __forceinline__ void put(int slot, double v) const { sm[slot * NT + q] = v; }
    __device__ __forceinline__ double get(int slot, int qq) const
    {
        return qq < 0 ? 0.0 : sm[slot * NT + qq];
    }
};

// which chunk of the line thread-chunk t of segment `seg` is, and whether it is stored
struct SegChunk {
    int chunk;       // chunk index on the periodic line
    bool interior;
};
__device__ __forceinline__ SegChunk seg_chunk(const SegGeom &sg, int seg, int t)
{
    if (sg.nseg == 1) return {t, true};
    const int first = seg * sg.iseg;                       // first interior chunk of the segment
    int c = first - sg.hlo + t;
    const bool interior = t >= sg.hlo && t < sg.hlo + sg.iseg && c < sg.NC;
    c %= sg.NC;
    if (c < 0) c += sg.NC;
    return {c, interior};
}

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

// CTA-wide barrier among the NT compute threads only (named barrier 1), so that a kernel may
// carry extra producer warps that do not take part
struct BarCompute {
    __device__ __forceinline__ void operator()() const { ptx::named_bar_sync_const<1, NT>(); }
};
struct BarAll {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// Solve NF right-hand sides in registers: v[f] <- A_f^-2 v[f].  Uses exchange slots
// [s0, s0 + 4*NF); two barriers.  Every compute thread of the CTA must call it.
template <int NF, class Bar>
__device__ __forceinline__ void solve_chunks(const CompositeCoef *const (&c)[NF], const Xchg &x,
                                             int s0, double (&v)[NF][LC], Bar bar)
{
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double ey, ez;
        fwd_local(c[f]->r, v[f], ey, ez);
        x.put(s0 + 2 * f, ey);
        x.put(s0 + 2 * f + 1, ez);
    }
    bar();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double Y, Z, ew, ex;
        lookback(*c[f], x, s0 + 2 * f, s0 + 2 * f + 1, -1, Y, Z);
        fwd_fix(*c[f], v[f], Y, Z);
        bwd_local(c[f]->r, v[f], ew, ex);
        x.put(s0 + 2 * NF + 2 * f, ew);
        x.put(s0 + 2 * NF + 2 * f + 1, ex);
    }
    bar();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double W, X;
        lookback(*c[f], x, s0 + 2 * NF + 2 * f, s0 + 2 * NF + 2 * f + 1, +1, W, X);
        bwd_fix(*c[f], v[f], W, X);
    }
}

// ---- the three pass bodies.  Inputs are the thread's chunk(s) in registers plus, for the field
// the derivative stencil acts on, its 3-point halos (e = {halo, chunk, halo}).  Exchange slots
// [0, nslots) are used; the caller guarantees that nobody still reads them from a previous tile.

// x pass:  A = Dxx f , B = Mxx f.   slots: 8 (solve) + 6 (halo of the solved B) = 14
constexpr int X_SLOTS = 14;
template <class Bar>
__device__ __forceinline__ void xpass_body(const CompositeCoef &M, const CompositeCoef &D,
                                           const Xchg &xc, const double (&ef)[LC + 6],
                                           double (&A)[LC], double (&B)[LC], Bar bar)
{
    double v[2][LC];
    stencil<true>(D, ef, v[0]);
#pragma unroll
    for (int k = 0; k < LC; ++k) v[1][k] = ef[k + 3];
    const CompositeCoef *const cs[2] = {&D, &M};
    solve_chunks<2>(cs, xc, 0, v, bar);
    put_halo(xc, 8, v[1]);
    bar();
    double e[LC + 6];
    get_halo(xc, 8, v[1], e);
    stencil<false>(M, e, B);
#pragma unroll
    for (int k = 0; k < LC; ++k) A[k] = v[0][k];
}

// y pass:  C = Myy a + Dyy b , Dd = Myy b.   slots: 12 (solve) + 12 (halos) = 24
constexpr int Y_SLOTS = 24;
template <class Bar>
__device__ __forceinline__ void ypass_body(const CompositeCoef &M, const CompositeCoef &D,
                                           const Xchg &xc, const double (&a)[LC],
                                           const double (&eb)[LC + 6], double (&C)[LC],
                                           double (&Dd)[LC], Bar bar)
{
    double v[3][LC];   // v0 = a, v1 = S_D b, v2 = b
    stencil<true>(D, eb, v[1]);
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        v[0][k] = a[k];
        v[2][k] = eb[k + 3];
    }
    const CompositeCoef *const cs[3] = {&M, &D, &M};
    solve_chunks<3>(cs, xc, 0, v, bar);
    put_halo(xc, 12, v[0]);
    put_halo(xc, 18, v[2]);
    bar();
    double e[LC + 6];
    get_halo(xc, 12, v[0], e);
    stencil<false>(M, e, C);
#pragma unroll
    for (int k = 0; k < LC; ++k) C[k] += v[1][k];
    get_halo(xc, 18, v[2], e);
    stencil<false>(M, e, Dd);
}

// z pass:  out = Mzz c + Dzz d.   slots: 8 (solve) + 6 (halo) = 14
constexpr int Z_SLOTS = 14;
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `early` runs just before the last barrier: the caller's loads issued there (the rows of p for the fused dot) have
// the barrier, the halo exchange and the last stencil to land, instead of a DRAM round trip after the last barrier
template <class Bar, class Hook = NoHook>
__device__ __forceinline__ void zpass_body(const CompositeCoef &M, const CompositeCoef &D,
                                           const Xchg &xc, const double (&c)[LC],
                                           const double (&ed)[LC + 6], double (&out)[LC], Bar bar,
                                           Hook early = Hook())
{
    double v[2][LC];   // v0 = c, v1 = S_D d
    stencil<true>(D, ed, v[1]);
#pragma unroll
    for (int k = 0; k < LC; ++k) v[0][k] = c[k];
    const CompositeCoef *const cs[2] = {&M, &D};
    solve_chunks<2>(cs, xc, 0, v, bar);
    put_halo(xc, 8, v[0]);
    early();
    bar();
    double e[LC + 6];
    get_halo(xc, 8, v[0], e);
    stencil<false>(M, e, out);
#pragma unroll
    for (int k = 0; k < LC; ++k) out[k] += v[1][k];
}

// ---- z pass of one slab of a z-decomposed box (ZOpen.open == 1) --------------------------------
// Same computation as zpass_body on an open line, with the neighbours' influence entering as
// (i) true stencil halos of the derivative input, (ii) a VIRTUAL chunk -1 / T whose published
// "end state" is the true recursion state at the slab boundary (the look-back then propagates it
// exactly: S_t = E_(t-1) + Phi S_(t-1)), (iii) true halos of the solved interpolation values.
// Chunk 0 and chunk T-1 of each line assemble these from the received numbers (ZOpen) and
// publish the virtual states in slots [ZV, ZV + 8).   slots: 14 + 8 = 22
constexpr int ZV = 14;
constexpr int ZNAT_SLOTS = 22;

// look-back that knows the virtual chunk: published states in slots (sy, sz), virtual ones in
// (vy, vz) of the thread that owns chunk 0 (dir = -1) or chunk T-1 (dir = +1) of the same line
__device__ __forceinline__ void lookback_nat(const CompositeCoef &c, const Xchg &x, int sy, int sz,
                                             int vy, int vz, int dir, double &Y, double &Z)
{
    Y = 0.0;
    Z = 0.0;
    const int qv = x.q + ((dir < 0 ? 0 : x.T - 1) - x.t) * x.tstride;
#pragma unroll
    for (int m = 1; m <= MAXLOOK; ++m) {
        if (m <= c.nlook) {
            const int tt = x.t + dir * m;
            double ey = 0.0, ez = 0.0;
            if (tt >= 0 && tt < x.T) {
                const int qm = x.q + (tt - x.t) * x.tstride;
                ey = x.sm[sy * NT + qm];
                ez = x.sm[sz * NT + qm];
            } else if (tt == -1 || tt == x.T) {
                ey = x.sm[vy * NT + qv];
                ez = x.sm[vz * NT + qv];
            }
            if (m == 1) {
                Y = ey;
                Z = ez;
            } else {
                const double p = c.look[m - 1];
                Y = fma(p, ey, Y);
                Z = fma(p, fma((double)(LC * (m - 1)), ey, ez), Z);
            }
        }
    }
}

// the neighbours' messages of one z line; issued early by the boundary chunks so that the loads
// overlap the wait for the tile instead of sitting in front of the first barrier
__device__ __forceinline__ void slab_load_messages(const ZOpen &zo, bool first, bool last,
                                                   long long line, double (&lo9)[DIST_MSG],
                                                   double (&up9)[DIST_MSG])
{
#pragma unroll
    for (int a = 0; a < DIST_MSG; ++a) {
        lo9[a] = first ? __ldg(zo.from_lo + a * zo.nlines + line) : 0.0;
        up9[a] = last ? __ldg(zo.from_up + a * zo.nlines + line) : 0.0;
    }
}

// ... as ONE array: a thread is chunk 0 or chunk T-1 of its line, never both (slabs hold at least four chunks), so
// the message it does not need costs no registers
__device__ __forceinline__ void slab_load_message(const ZOpen &zo, bool first, bool last, long long line,
                                                  double (&m9)[DIST_MSG])
{
    const double *src = first ? zo.from_lo : zo.from_up;
#pragma unroll
    for (int a = 0; a < DIST_MSG; ++a) m9[a] = (first || last) ? __ldg(src + a * zo.nlines + line) : 0.0;
}

template <class Bar, class Hook = NoHook>
__device__ __forceinline__ void zpass_body_slab(const CompositeCoef &M, const CompositeCoef &D,
                                                const ZOpen &zo, const Xchg &xc,
                                                const double (&lo9)[DIST_MSG],
                                                const double (&up9)[DIST_MSG],
                                                const double (&c)[LC], double (&ed)[LC + 6],
                                                double (&out)[LC], Bar bar, Hook early = Hook())
{
    const bool first = xc.t == 0, last = xc.t == xc.T - 1;
    // (i) true halos of the derivative input
    if (first) {
        ed[0] = lo9[8];
        ed[1] = lo9[7];
        ed[2] = lo9[6];
    }
    if (last) {
        ed[LC + 3] = up9[4];
        ed[LC + 4] = up9[5];
        ed[LC + 5] = up9[6];
    }
    double v[2][LC];   // v0 = c (interpolation, solve first), v1 = S_D d (derivative, stencil first)
    stencil<true>(D, ed, v[1]);
#pragma unroll
    for (int k = 0; k < LC; ++k) v[0][k] = c[k];

    // (ii) virtual causal states at plane -1, published by chunk 0
    if (first) {
        const double d0 = ed[3], d1 = ed[4], d2 = ed[5], r = D.r;
        // the lower rank's derivative recursion ran on a zero-halo stencil: add what my first
        // three planes contribute to its last three right-hand sides
        const double s1 = fma(D.c3, d2, fma(D.c2, d1, D.c1 * d0));
        const double s2 = fma(D.c3, d1, D.c2 * d0);
        const double s3 = D.c3 * d0;
        xc.put(ZV + 0, lo9[0]);
        xc.put(ZV + 1, lo9[1]);
        xc.put(ZV + 2, lo9[4] + fma(r, fma(r, s3, s2), s1));
        xc.put(ZV + 3, lo9[5] + fma(r, fma(3.0 * r, s3, 2.0 * s2), s1));
    }
    double ey[2], ez[2];
    fwd_local(M.r, v[0], ey[0], ez[0]);
    fwd_local(D.r, v[1], ey[1], ez[1]);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        xc.put(2 * f, ey[f]);
        xc.put(2 * f + 1, ez[f]);
    }
    bar();
    double WT = 0.0, XT = 0.0, yMe = 0.0, zMe = 0.0;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        const CompositeCoef &cf = f == 0 ? M : D;
        double Y, Z, ew, ex;
        lookback_nat(cf, xc, 2 * f, 2 * f + 1, ZV + 2 * f, ZV + 2 * f + 1, -1, Y, Z);
        fwd_fix(cf, v[f], Y, Z);
        if (last) {
            // my true outgoing causal state, and with it the true anti-causal state at plane T*16
            const double yt = fma(cf.pw[LC - 1], Y, ey[f]), zt = v[f][LC - 1];
            double a0 = up9[2 * f], a1 = up9[2 * f + 1];
            if (f == 1) {
                // my last three planes contribute to the upper rank's first three right-hand sides
                const double dm1 = ed[LC + 2], dm2 = ed[LC + 1], dm3 = ed[LC], r = D.r;
                const double e0 = fma(D.c3, dm3, fma(D.c2, dm2, D.c1 * dm1));
                const double e1 = fma(D.c3, dm2, D.c2 * dm1);
                const double e2 = D.c3 * dm1;
                a0 += fma(r, fma(r, e2, e1), e0);
                a1 += r * fma(2.0 * r, e2, e1);
            }
            const double W = fma(zo.gw[f], a0, fma(zo.kwz[f], zt, zo.kwy[f] * yt));
            const double X = fma(zo.gx0[f], a0, fma(zo.gw[f], a1, fma(zo.kxz[f], zt, zo.kxy[f] * yt)));
            xc.put(ZV + 4 + 2 * f, W);
            xc.put(ZV + 5 + 2 * f, X);
            if (f == 0) {
                WT = W;
                XT = X;
                yMe = yt;
                zMe = zt;
            }
        }
        bwd_local(cf.r, v[f], ew, ex);
        xc.put(4 + 2 * f, ew);
        xc.put(5 + 2 * f, ex);
    }
    bar();
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        const CompositeCoef &cf = f == 0 ? M : D;
        double W, X;
        lookback_nat(cf, xc, 4 + 2 * f, 5 + 2 * f, ZV + 4 + 2 * f, ZV + 5 + 2 * f, +1, W, X);
        bwd_fix(cf, v[f], W, X);
    }
    // (iii) halos of the solved interpolation values
    put_halo(xc, 8, v[0]);
    double hl[3] = {0, 0, 0}, hh[3] = {0, 0, 0};
    if (first) {
        // continue the anti-causal recursion into the lower rank's top three planes
        const double r = M.r;
        double x = v[0][0], w = fma(-r, v[0][1], x);
        w = fma(r, w, lo9[1]);
        x = fma(r, x, w);
        hl[2] = x;
        w = fma(r, w, lo9[2]);
        x = fma(r, x, w);
        hl[1] = x;
        w = fma(r, w, lo9[3]);
        x = fma(r, x, w);
        hl[0] = x;
    }
    if (last) {
        // x at the first plane above is the virtual state itself; two more planes by running the
        // recursion upwards, with the true causal values there (upper rank's raw c0, c1 plus the
        // homogeneous solution of my outgoing state)
        const double r = M.r, ri = zo.rinv;
        const double zt0 = fma(r, zMe + yMe, up9[7]);
        const double zt1 = fma(r * r, fma(2.0, yMe, zMe), fma(2.0 * r, up9[7], up9[8]));
        const double w1 = (WT - zt0) * ri, x1 = (XT - WT) * ri;
        const double w2 = (w1 - zt1) * ri, x2 = (x1 - w1) * ri;
        (void)w2;
        hh[0] = XT;
        hh[1] = x1;
        hh[2] = x2;
    }
    early();
    bar();
    double e[LC + 6];
    get_halo(xc, 8, v[0], e);
    if (first) {
        e[0] = hl[0];
        e[1] = hl[1];
        e[2] = hl[2];
    }
    if (last) {
        e[LC + 3] = hh[0];
        e[LC + 4] = hh[1];
        e[LC + 5] = hh[2];
    }
    stencil<false>(M, e, out);
#pragma unroll
    for (int k = 0; k < LC; ++k) out[k] += v[1][k];
}

// Deterministic sum over `nthr` compute threads (q = 0 .. nthr-1), result valid in thread q == 0.
// Full warps: fixed shuffle tree per warp, warp sums through `ws` (>= 32 doubles that nobody else
// touches between two calls), ONE barrier.  Otherwise the generic fixed-order version below.
template <class Bar>
__device__ __forceinline__ double block_sum_warps(double v, double *ws, int q, int nthr, Bar bar)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((q & 31) == 0) ws[q >> 5] = v;
    bar();
    double tot = 0.0;
    if (q == 0) {
        const int nw = nthr >> 5;
        for (int i = 0; i < nw; ++i) tot += ws[i];
    }
    return tot;
}

// fixed-shape (deterministic) sum over the compute threads of a CTA through the exchange area;
// result valid in thread q == 0.  Two barriers before the area may be reused.
template <class Bar>
__device__ __forceinline__ double block_sum_fixed(double v, double *sm, int q, int nthr, Bar bar)
{
    bar();
    sm[q] = v;
    bar();
    double s = 0.0;
    if (q < 32)
        for (int i = q; i < nthr; i += 32) s += sm[i];
    bar();
    if (q < 32) sm[q] = s;
    bar();
    double tot = 0.0;
    if (q == 0) {
        const int m = nthr < 32 ? nthr : 32;
        for (int i = 0; i < m; ++i) tot += sm[i];
    }
    return tot;
}

}  // namespace fast
}  // namespace pbx
