// pbx_cg_dev.cuh -- device-side pieces of the conjugate-gradient loop (pbx_cg.cu) that other kernels
// use too: the layout of the scalar block, the fixed-shape CTA sum, the scalar step of the loop, and
// the reduction TAIL that lets the kernel which wrote the per-CTA partial sums finish the job itself.
#pragma once

#include "pbx_internal.h"
#include "pbx_peer.cuh"

namespace pbx {
namespace cgdev {

enum {
    SC_M0 = 0, SC_S1, SC_S2, SC_PW, SC_BETA, SC_BETAOLD, SC_A, SC_B, SC_DP, SC_DP0, SC_TTOL,
    SC_PWOLD, SC_STATUS, SC_IT, SC_RTOL, SC_ABSTOL, SC_MEAN, SC_MAXIT, SC_NTOT,
    // preconditioned CG: sums of z, z^2, z (r - m0), (r - m0) (contiguous: one reduction), mean of z
    SC_SZ, SC_SZZ, SC_SZR, SC_SR, SC_MZ,
    SC_XIT,   // number of the iteration whose step length SC_A is (x += a p of that iteration pending)
    SC_COUNT
};

constexpr int VT = 256;

// fixed-shape sum over the NTHREADS threads of a CTA: warp shuffle tree, then warp 0 over the warp
// sums; result valid in thread 0.  sh: NTHREADS / 32 doubles of shared memory.
template <int NTHREADS>
__device__ __forceinline__ double block_sum_n(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < NTHREADS / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    }
    return s;
}
__device__ __forceinline__ double block_sum(double v, double *sh) { return block_sum_n<VT>(v, sh); }

// the scalar logic of the KSPCG loop; one thread
//   phase 0: m0 = S1 / N                     (S1 = sum b)
//   phase 1: initial residual norm / test    (S1, S2 about m0)
//   phase 2: a = beta / (p.w), indefiniteness test
//   phase 3: new residual norm, test, b = beta/beta_old
// with a preconditioner (z = M^-1 r, sums SC_SZ .. SC_SR of z, z^2, z (r - m0), r - m0):
//   phase 4: first application: mean of z, ||z||, beta = z.r, test
//   phase 5: mean of the updated residual (S1 about m0), the right-hand side's mean for the next PC
//   phase 6: as phase 4 after an iteration: counts it, b = beta/beta_old, tests
//   phase 7 / 8: SC_MEAN / SC_MZ = S1 / N (stand-alone preconditioner application)
__device__ __forceinline__ void scalar_phase(double *__restrict__ sc, int phase, double *__restrict__ hist, int nhist)
{
    const double N = sc[SC_NTOT];
    if (phase == 0) {
        sc[SC_M0] = sc[SC_S1] / N;
        sc[SC_MEAN] = sc[SC_M0];
        return;
    }
    if (phase == 7 || phase == 8) {
        sc[phase == 7 ? SC_MEAN : SC_MZ] = sc[SC_S1] / N;
        return;
    }
    if (sc[SC_STATUS] != 0.0) return;
    if (phase == 5) {
        sc[SC_MEAN] = sc[SC_M0] + sc[SC_S1] / N;
        return;
    }
    if (phase == 4 || phase == 6) {
        const double mz = sc[SC_SZ] / N;
        double zz = sc[SC_SZZ] - sc[SC_SZ] * mz;
        if (zz < 0.0) zz = 0.0;
        const double dp = sqrt(zz);
        const double beta = sc[SC_SZR] - mz * sc[SC_SR];   // (z - mz) . r
        sc[SC_MZ] = mz;
        sc[SC_DP] = dp;
        int it = (int)sc[SC_IT];
        if (phase == 4) {
            sc[SC_DP0] = dp;
            sc[SC_TTOL] = fmax(sc[SC_RTOL] * dp, sc[SC_ABSTOL]);
            sc[SC_BETA] = beta;
            sc[SC_B] = 0.0;
            sc[SC_PWOLD] = 0.0;
            if (hist && nhist > 0) hist[0] = dp;
        } else {
            it += 1;
            sc[SC_IT] = it;
            sc[SC_BETAOLD] = sc[SC_BETA];
            sc[SC_BETA] = beta;
            sc[SC_B] = beta / sc[SC_BETAOLD];
            if (hist && it < nhist) hist[it] = dp;
        }
        if (dp != dp || beta != beta)
            sc[SC_STATUS] = PBX_DIVERGED_NANORINF;
        else if (dp <= sc[SC_TTOL])
            sc[SC_STATUS] = dp < sc[SC_ABSTOL] ? PBX_CONVERGED_ATOL : PBX_CONVERGED_RTOL;
        else if (beta < 0.0)
            sc[SC_STATUS] = PBX_DIVERGED_INDEFINITE_PC;
        else if (phase == 6 && dp >= 1.0e4 * sc[SC_DP0])
            sc[SC_STATUS] = PBX_DIVERGED_DTOL;
        else if (phase == 6 && it >= (int)sc[SC_MAXIT])
            sc[SC_STATUS] = PBX_DIVERGED_ITS;
        return;
    }
    if (phase == 2) {
        const double dpi = sc[SC_PW], dpiold = sc[SC_PWOLD];
        const int i = (int)sc[SC_IT];
        const double beta = sc[SC_BETA];
        if (beta == 0.0) {
            sc[SC_IT] = i + 1;
            sc[SC_STATUS] = PBX_CONVERGED_ATOL;
            return;
        }
        if (dpi != dpi) {
            sc[SC_IT] = i + 1;
            sc[SC_STATUS] = PBX_DIVERGED_NANORINF;
            return;
        }
        const double sg = (dpi > 0) - (dpi < 0), sgo = (dpiold > 0) - (dpiold < 0);
        if (dpi == 0.0 || (i > 0 && sg * sgo < 0.0)) {
            sc[SC_IT] = i + 1;
            sc[SC_STATUS] = PBX_DIVERGED_INDEFINITE_MAT;
            return;
        }
        sc[SC_PWOLD] = dpi;
        sc[SC_A] = beta / dpi;
        sc[SC_XIT] = i + 1;
        return;
    }
    // phases 1 and 3: S1, S2 are sums of (r - m0), (r - m0)^2
    const double dm = sc[SC_S1] / N;
    double zz = sc[SC_S2] - sc[SC_S1] * dm;
    if (zz < 0.0) zz = 0.0;
    const double dp = sqrt(zz);
    sc[SC_MEAN] = sc[SC_M0] + dm;
    sc[SC_DP] = dp;
    int it = (int)sc[SC_IT];
    if (phase == 1) {
        sc[SC_DP0] = dp;
        sc[SC_TTOL] = fmax(sc[SC_RTOL] * dp, sc[SC_ABSTOL]);
        sc[SC_BETA] = zz;
        sc[SC_PWOLD] = 0.0;
        if (hist && nhist > 0) hist[0] = dp;
    } else {
        it += 1;
        sc[SC_IT] = it;
        sc[SC_BETAOLD] = sc[SC_BETA];
        sc[SC_BETA] = zz;
        sc[SC_B] = zz / sc[SC_BETAOLD];
        if (hist && it < nhist) hist[it] = dp;
    }
    if (dp != dp)
        sc[SC_STATUS] = PBX_DIVERGED_NANORINF;
    else if (dp <= sc[SC_TTOL])
        sc[SC_STATUS] = dp < sc[SC_ABSTOL] ? PBX_CONVERGED_ATOL : PBX_CONVERGED_RTOL;
    else if (phase == 3 && dp >= 1.0e4 * sc[SC_DP0])
        sc[SC_STATUS] = PBX_DIVERGED_DTOL;
    else if (phase == 3 && it >= (int)sc[SC_MAXIT])
        sc[SC_STATUS] = PBX_DIVERGED_ITS;
}


// The reduction tail (RedTail, pbx_internal.h; PBX_FUSE_TAIL=0 turns it off).  Called by ALL NTHREADS threads
// of EVERY CTA at the end of a kernel that has written its per-CTA partial sums: the CTA that draws the
// last ticket sums the partials in a fixed shape, all-reduces them over the peer boards (several
// ranks), stores them into the scalar block and runs the scalar step of the loop -- what k_reduce (+
// exchange) + k_scalar do in further launches.  The ticket counter is left at zero for the next launch.
template <int NTHREADS>
__device__ __forceinline__ void red_tail(const RedTail &t)
{
    __shared__ double sh[NTHREADS / 32];
    __shared__ double mine[PEER_VALS];
    __shared__ double all[PEER_MAXR][PEER_VALS];
    __shared__ int s_last;
    __threadfence();                               // my partial sums are visible before my ticket is drawn
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(t.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (!(t.guarded && t.sc[SC_STATUS] != 0.0)) {  // the same verdict on every rank
        for (int a = 0; a < t.narr; ++a) {
            // (fixed shape, but not k_reduce's: NTHREADS strides instead of VT -- a sum taken here and one taken
            // in a launch of its own agree to rounding, not bit for bit)
            double s = 0.0;
            for (int i = threadIdx.x; i < t.cnt; i += NTHREADS) s += t.part[a * t.stride + i];
            s = block_sum_n<NTHREADS>(s, sh);
            if (threadIdx.x == 0) mine[a] = s;
            __syncthreads();
        }
        double res[PEER_VALS];
        if (t.L.n > 1) {
            peer_exchange_sum(t.L, t.seq, mine, t.narr, all, res);
        } else if (threadIdx.x == 0) {
            for (int a = 0; a < t.narr; ++a) res[a] = mine[a];
        }
        if (threadIdx.x == 0) {
            for (int a = 0; a < t.narr; ++a) t.dst[a] = res[a];
            if (t.phase >= 0) scalar_phase(t.sc, t.phase, t.hist, t.nhist);
        }
    }
    if (threadIdx.x == 0) *t.ticket = 0u;
}

}  // namespace cgdev
}  // namespace pbx
