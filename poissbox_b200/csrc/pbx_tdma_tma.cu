// pbx_tdma_tma.cu -- batched general-coefficient tridiagonal solves for LINE-MAJOR batches
// (element stride 1: point i of line l at base[l*ls + i], the layout of the reference's own calls
// `tdma(a(:), b(:), c(:), d(:))` on contiguous lines, src/tridsol.f90:22-115), sm_100a.
//
// The arithmetic is that of pbx_tdma.cu -- one thread per line, the reference's operations in the
// reference's order with round-to-nearest intrinsics, hence the same bits as the CPU oracle.  What
// changes is the data movement.  With contiguous lines a thread-per-line kernel that loads from
// global memory itself touches one 32-byte sector per 8-byte load and 32 different 128-byte lines
// per warp instruction; here the lines travel as TMA tiles instead:
//
//   * a CTA is ONE warp = 32 lines; a tile is 16 points x 32 lines of one array (4 KiB: a 128-byte
//     row per line), fetched by cp.async.bulk.tensor with the 128-byte swizzle, so that lane l reads
//     its row with conflict-free 128-bit shared loads (the x pass's pattern, pbx_fast_tma.cu);
//   * two stages of four tiles (a, b, c, d); the results (pivots b', d'; the solution) are written
//     IN PLACE into the tiles they came from and leave by TMA stores from there, so a thread owns
//     1 KiB of shared memory and ~7 CTAs = 220 lines are resident per SM -- what the serial divide
//     chain (one IEEE division per point and sweep) needs to cover its latency;
//   * lines that do not fill the last block or the last CTA need no special code: the tensor maps
//     carry the true extents, out-of-range loads are zero-filled and out-of-range stores dropped.
//
// Needs es == 1, an even n >= 2, an even line stride and 16-byte aligned bases; everything else stays
// on the generic kernels.
#include <cuda.h>

#include <cstdlib>

#include "pbx_internal.h"
#include "pbx_ptx.cuh"

namespace pbx {

namespace {

using namespace ptx;

constexpr int LM_LINES = 32;                  // lines per CTA (one warp, a thread per line)
constexpr int LM_PTS = 16;                    // points per block: one 128-byte row per line
constexpr int LM_TILE = LM_LINES * LM_PTS;    // doubles per array tile (4 KiB)
constexpr uint32_t LM_TILE_BYTES = LM_TILE * 8;
constexpr int LM_STAGES = 2;

template <int NARR>
struct LmSharedT {
    double t[LM_STAGES][NARR][LM_TILE];       // [stage][array][row * 16 + swizzled piece]
    uint64_t full[LM_STAGES];
};
using LmShared = LmSharedT<4>;                // 32 KiB: six CTAs = 192 lines per SM
using LmShared3 = LmSharedT<3>;               // backward sweep: 24 KiB, nine CTAs per SM

// 16-byte piece j of row q under the 128-byte swizzle (tiles are 1 KiB aligned)
__device__ __forceinline__ int swz(int q, int j) { return q * 16 + ((j ^ (q & 7)) << 1); }

__device__ __forceinline__ void row_read8(const double *tile, int lane, int half, double (&v)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2 x = *reinterpret_cast<const double2 *>(tile + swz(lane, half * 4 + j));
        v[2 * j] = x.x;
        v[2 * j + 1] = x.y;
    }
}
__device__ __forceinline__ void row_write8(double *tile, int lane, int half, const double (&v)[8])
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2 *>(tile + swz(lane, half * 4 + j)) = make_double2(v[2 * j], v[2 * j + 1]);
}

struct LmMaps {
    CUtensorMap m[4];
};

// lane 0: fetch block `blk` (points 16 blk ..) of `narr` arrays into stage s
template <int NARR>
__device__ __forceinline__ void lm_issue(LmSharedT<NARR> &S, const LmMaps &M, int s, int blk, int line0)
{
    mbar_expect_tx(&S.full[s], (uint32_t)NARR * LM_TILE_BYTES);
#pragma unroll
    for (int q = 0; q < NARR; ++q)
        tma_load_2d(S.t[s][q], &M.m[4 - NARR + q], &S.full[s], blk * LM_PTS, line0);
}

// fwd_sweep, src/tridsol.f90:76-96, maps = (a, b, c, d); b and d leave as pivots / reduced rhs
__global__ void __launch_bounds__(LM_LINES)
fwd_lm_kernel(int n, const __grid_constant__ LmMaps M)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    LmShared &S = *reinterpret_cast<LmShared *>(smraw);
    if ((smem_u32(smraw) & 1023u) != 0) __trap();
    const int lane = threadIdx.x, line0 = blockIdx.x * LM_LINES;
    const int nblk = (n + LM_PTS - 1) / LM_PTS;
    if (lane == 0) {
        for (int s = 0; s < LM_STAGES; ++s) mbar_init(&S.full[s], 1);
        fence_mbar_init();
        for (int s = 0; s < LM_STAGES && s < nblk; ++s) lm_issue(S, M, s, s, line0);
    }
    __syncwarp();
    double bp = 0.0, dp = 0.0, cp = 0.0;
    for (int k = 0; k < nblk; ++k) {
        const int s = k & 1;
        mbar_wait(&S.full[s], (uint32_t)((k >> 1) & 1));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double a[8], b[8], c[8], d[8];
            row_read8(S.t[s][0], lane, half, a);
            row_read8(S.t[s][1], lane, half, b);
            row_read8(S.t[s][2], lane, half, c);
            row_read8(S.t[s][3], lane, half, d);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = k * LM_PTS + half * 8 + u;
                if (i == 0) {
                    bp = b[u];
                    dp = d[u];
                    cp = c[u];
                } else if (i < n) {
                    const double w = __ddiv_rn(a[u], bp);                 // :91
                    bp = __dsub_rn(b[u], __dmul_rn(w, cp));               // :92
                    dp = __dsub_rn(d[u], __dmul_rn(w, dp));               // :93
                    cp = c[u];
                    b[u] = bp;
                    d[u] = dp;
                }
            }
            row_write8(S.t[s][1], lane, half, b);
            row_write8(S.t[s][3], lane, half, d);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&M.m[1], S.t[s][1], k * LM_PTS, line0);
            tma_store_2d(&M.m[3], S.t[s][3], k * LM_PTS, line0);
            tma_commit();
            if (k + LM_STAGES < nblk) {
                tma_wait_read0();     // the stores have read the stage: it can be refilled
                lm_issue(S, M, s, k + LM_STAGES, line0);
            }
        }
    }
    if (lane == 0) tma_wait_all0();
}

// bwd_sweep, src/tridsol.f90:98-115, maps = (-, b, c, d); d leaves as the solution
__global__ void __launch_bounds__(LM_LINES)
bwd_lm_kernel(int n, const __grid_constant__ LmMaps M)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    LmShared3 &S = *reinterpret_cast<LmShared3 *>(smraw);
    if ((smem_u32(smraw) & 1023u) != 0) __trap();
    const int lane = threadIdx.x, line0 = blockIdx.x * LM_LINES;
    const int nblk = (n + LM_PTS - 1) / LM_PTS;
    if (lane == 0) {
        for (int s = 0; s < LM_STAGES; ++s) mbar_init(&S.full[s], 1);
        fence_mbar_init();
        for (int s = 0; s < LM_STAGES && s < nblk; ++s) lm_issue(S, M, s, nblk - 1 - s, line0);
    }
    __syncwarp();
    double x = 0.0;
    for (int kk = 0; kk < nblk; ++kk) {
        const int s = kk & 1, k = nblk - 1 - kk;
        mbar_wait(&S.full[s], (uint32_t)((kk >> 1) & 1));
#pragma unroll
        for (int half = 1; half >= 0; --half) {
            double b[8], c[8], d[8];
            row_read8(S.t[s][0], lane, half, b);
            row_read8(S.t[s][1], lane, half, c);
            row_read8(S.t[s][2], lane, half, d);
#pragma unroll
            for (int u = 7; u >= 0; --u) {
                const int i = k * LM_PTS + half * 8 + u;
                if (i == n - 1) {
                    x = __ddiv_rn(d[u], b[u]);                                        // :108
                    d[u] = x;
                } else if (i < n) {
                    x = __ddiv_rn(__dsub_rn(d[u], __dmul_rn(c[u], x)), b[u]);         // :111
                    d[u] = x;
                }
            }
            row_write8(S.t[s][2], lane, half, d);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&M.m[3], S.t[s][2], k * LM_PTS, line0);
            tma_commit();
            if (kk + LM_STAGES < nblk) {
                tma_wait_read0();
                lm_issue(S, M, s, k - LM_STAGES, line0);
            }
        }
    }
    if (lane == 0) tma_wait_all0();
}

// tdma_periodic, src/tridsol.f90:34-74: the two Thomas solves of the Sherman-Morrison closure (on d
// and on u = (gamma, 0 .. 0, c(n))) share the pivots and run together; maps = (a, b, c, d) plus the
// two workspace arrays (modified pivots, u), line-major like the data.  Three legs in one kernel:
//   forward   loads a, b, c, d      stores bmod' (from b's tile), d', u' (from a's tile)
//   backward  loads bmod', c, d', u'  stores the two solutions (d's and u's tiles)
//   combine   loads y, q              stores d = y - q fac / den                          (:69-70)
struct LmMapsP {
    CUtensorMap a, b, c, d, wb, wu;
};

__global__ void __launch_bounds__(LM_LINES)
periodic_lm_kernel(int n, long long nl, long long ls, const __grid_constant__ LmMapsP M,
                   const double *__restrict__ ga, const double *__restrict__ gb,
                   const double *__restrict__ gc)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    LmShared &S = *reinterpret_cast<LmShared *>(smraw);
    if ((smem_u32(smraw) & 1023u) != 0) __trap();
    const int lane = threadIdx.x, line0 = blockIdx.x * LM_LINES;
    const int nblk = (n + LM_PTS - 1) / LM_PTS;
    // the closure's scalars (:51-56); lanes beyond the batch work on zeros and store nothing
    const long long l = (long long)line0 + lane;
    const bool live = l < nl;
    const long long o = live ? l * ls : 0;
    const double b1 = live ? gb[o] : 1.0, a1 = live ? ga[o] : 0.0, cn_ = live ? gc[o + n - 1] : 0.0;
    const double bn = live ? gb[o + n - 1] : 1.0;
    const double gamma = -b1;                                                         // :51
    const double b1m = __dsub_rn(b1, gamma);                                          // :55
    const double bnm = __dsub_rn(bn, __ddiv_rn(__dmul_rn(cn_, a1), gamma));           // :56

    if (lane == 0) {
        for (int s = 0; s < LM_STAGES; ++s) mbar_init(&S.full[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    unsigned ph = 0;   // bit s: phase parity of stage s's barrier

    // ---- forward leg
    auto issue_fwd = [&](int s, int blk) {
        mbar_expect_tx(&S.full[s], 4 * LM_TILE_BYTES);
        tma_load_2d(S.t[s][0], &M.a, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][1], &M.b, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][2], &M.c, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][3], &M.d, &S.full[s], blk * LM_PTS, line0);
    };
    if (lane == 0)
        for (int s = 0; s < LM_STAGES && s < nblk; ++s) issue_fwd(s, s);
    double bp = 0.0, dp = 0.0, up = 0.0, cp = 0.0;
    for (int k = 0; k < nblk; ++k) {
        const int s = k & 1;
        mbar_wait(&S.full[s], (ph >> s) & 1u);
        ph ^= 1u << s;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double a[8], b[8], c[8], d[8];
            row_read8(S.t[s][0], lane, half, a);
            row_read8(S.t[s][1], lane, half, b);
            row_read8(S.t[s][2], lane, half, c);
            row_read8(S.t[s][3], lane, half, d);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = k * LM_PTS + half * 8 + u;
                if (i == 0) {
                    bp = b1m;
                    dp = d[u];
                    up = gamma;                                                       // :63
                    cp = c[u];
                    b[u] = bp;
                    a[u] = up;
                } else if (i < n) {
                    const double bi = (i == n - 1) ? bnm : b[u];
                    const double ui = (i == n - 1) ? cn_ : 0.0;                       // :64-65
                    const double w = __ddiv_rn(a[u], bp);
                    bp = __dsub_rn(bi, __dmul_rn(w, cp));
                    dp = __dsub_rn(d[u], __dmul_rn(w, dp));
                    up = __dsub_rn(ui, __dmul_rn(w, up));
                    cp = c[u];
                    b[u] = bp;
                    d[u] = dp;
                    a[u] = up;
                }
            }
            row_write8(S.t[s][0], lane, half, a);
            row_write8(S.t[s][1], lane, half, b);
            row_write8(S.t[s][3], lane, half, d);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&M.wu, S.t[s][0], k * LM_PTS, line0);
            tma_store_2d(&M.wb, S.t[s][1], k * LM_PTS, line0);
            tma_store_2d(&M.d, S.t[s][3], k * LM_PTS, line0);
            tma_commit();
            if (k + LM_STAGES < nblk) {
                tma_wait_read0();
                issue_fwd(s, k + LM_STAGES);
            }
        }
    }
    // the backward leg reads what the forward leg stored: the stores must be complete (and the
    // stages free) before the first load is issued
    if (lane == 0) tma_wait_all0();
    __syncwarp();

    // ---- backward leg
    auto issue_bwd = [&](int s, int blk) {
        mbar_expect_tx(&S.full[s], 4 * LM_TILE_BYTES);
        tma_load_2d(S.t[s][0], &M.wu, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][1], &M.wb, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][2], &M.c, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][3], &M.d, &S.full[s], blk * LM_PTS, line0);
    };
    if (lane == 0)
        for (int s = 0; s < LM_STAGES && s < nblk; ++s) issue_bwd(s, nblk - 1 - s);
    double xd = 0.0, xu = 0.0, dn_ = 0.0, un = 0.0;
    for (int kk = 0; kk < nblk; ++kk) {
        const int s = kk & 1, k = nblk - 1 - kk;
        mbar_wait(&S.full[s], (ph >> s) & 1u);
        ph ^= 1u << s;
#pragma unroll
        for (int half = 1; half >= 0; --half) {
            double v[8], b[8], c[8], d[8];
            row_read8(S.t[s][0], lane, half, v);
            row_read8(S.t[s][1], lane, half, b);
            row_read8(S.t[s][2], lane, half, c);
            row_read8(S.t[s][3], lane, half, d);
#pragma unroll
            for (int u = 7; u >= 0; --u) {
                const int i = k * LM_PTS + half * 8 + u;
                if (i == n - 1) {
                    xd = __ddiv_rn(d[u], b[u]);
                    xu = __ddiv_rn(v[u], b[u]);
                    dn_ = xd;
                    un = xu;
                    d[u] = xd;
                    v[u] = xu;
                } else if (i < n) {
                    xd = __ddiv_rn(__dsub_rn(d[u], __dmul_rn(c[u], xd)), b[u]);
                    xu = __ddiv_rn(__dsub_rn(v[u], __dmul_rn(c[u], xu)), b[u]);
                    d[u] = xd;
                    v[u] = xu;
                }
            }
            row_write8(S.t[s][0], lane, half, v);
            row_write8(S.t[s][3], lane, half, d);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&M.wu, S.t[s][0], k * LM_PTS, line0);
            tma_store_2d(&M.d, S.t[s][3], k * LM_PTS, line0);
            tma_commit();
            if (kk + LM_STAGES < nblk) {
                tma_wait_read0();
                issue_bwd(s, k - LM_STAGES);
            }
        }
    }
    if (lane == 0) tma_wait_all0();
    __syncwarp();

    // ---- combine (:69-70): xd, xu now hold the solutions at point 0
    const double a1g = __ddiv_rn(a1, gamma);
    const double fac = __dadd_rn(xd, __dmul_rn(a1g, dn_));
    const double den = __dadd_rn(1.0, __dadd_rn(xu, __dmul_rn(a1g, un)));
    auto issue_cmb = [&](int s, int blk) {
        mbar_expect_tx(&S.full[s], 2 * LM_TILE_BYTES);
        tma_load_2d(S.t[s][0], &M.wu, &S.full[s], blk * LM_PTS, line0);
        tma_load_2d(S.t[s][3], &M.d, &S.full[s], blk * LM_PTS, line0);
    };
    if (lane == 0)
        for (int s = 0; s < LM_STAGES && s < nblk; ++s) issue_cmb(s, s);
    for (int k = 0; k < nblk; ++k) {
        const int s = k & 1;
        mbar_wait(&S.full[s], (ph >> s) & 1u);
        ph ^= 1u << s;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double v[8], d[8];
            row_read8(S.t[s][0], lane, half, v);
            row_read8(S.t[s][3], lane, half, d);
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] = __dsub_rn(d[u], __ddiv_rn(__dmul_rn(v[u], fac), den));
            row_write8(S.t[s][3], lane, half, d);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&M.d, S.t[s][3], k * LM_PTS, line0);
            tma_commit();
            if (k + LM_STAGES < nblk) {
                tma_wait_read0();
                issue_cmb(s, k + LM_STAGES);
            }
        }
    }
    if (lane == 0) tma_wait_all0();
}

// default since its first B200 run (line-major batches: tdma 20 -> 51 Gpt/s, tdma_periodic 10.5 -> 20.7 Gpt/s
// for 512-point lines); PBX_TDMA_TMA=0 selects the generic thread-per-line kernels
bool lm_enabled() { return env_switch("PBX_TDMA_TMA", true); }

bool lm_shape_ok(int n, long long nl, long long es, long long ls, const void *p0, const void *p1,
                 const void *p2, const void *p3)
{
    // n even: TMA moves 16-byte units, so the last unit of an odd line would reach one element past the
    // line (measured on the B200: the store wrote the zero fill into the caller's padding element)
    if (!lm_enabled() || es != 1 || n < 2 || (n & 1) || nl < 1 || ls < n || (ls & 1)) return false;
    if (nl > 0x7fffffffLL - LM_LINES) return false;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) |
                         reinterpret_cast<uintptr_t>(p2) | reinterpret_cast<uintptr_t>(p3);
    return (al & 15) == 0 && fast_tma_available();
}

bool lm_map(CUtensorMap *m, const double *base, int n, long long nl, long long ls)
{
    return tma_make_map_2d(m, base, (unsigned long long)n, (unsigned long long)nl,
                           (unsigned long long)ls * 8, LM_PTS, LM_LINES, true);
}

}  // namespace

// PBX_ERR_UNSUPPORTED: the caller runs the generic kernel
int tdma_fwd_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                      double *b, const double *c, double *d)
{
    if (!lm_shape_ok(n, nl, es, ls, a, b, c, d)) return PBX_ERR_UNSUPPORTED;
    LmMaps M;
    if (!lm_map(&M.m[0], a, n, nl, ls) || !lm_map(&M.m[1], b, n, nl, ls) || !lm_map(&M.m[2], c, n, nl, ls) ||
        !lm_map(&M.m[3], d, n, nl, ls))
        return PBX_ERR_UNSUPPORTED;
    fwd_lm_kernel<<<(unsigned)((nl + LM_LINES - 1) / LM_LINES), LM_LINES, sizeof(LmShared), s>>>(n, M);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int tdma_bwd_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *b,
                      const double *c, double *d)
{
    if (!lm_shape_ok(n, nl, es, ls, b, c, d, d)) return PBX_ERR_UNSUPPORTED;
    LmMaps M;
    if (!lm_map(&M.m[1], b, n, nl, ls) || !lm_map(&M.m[2], c, n, nl, ls) || !lm_map(&M.m[3], d, n, nl, ls))
        return PBX_ERR_UNSUPPORTED;
    M.m[0] = M.m[1];
    bwd_lm_kernel<<<(unsigned)((nl + LM_LINES - 1) / LM_LINES), LM_LINES, sizeof(LmShared3), s>>>(n, M);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// ws: 2 * nl * n doubles (16-byte aligned), line-major like the data
int tdma_periodic_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                           const double *b, const double *c, double *d, double *ws)
{
    if (!lm_shape_ok(n, nl, es, ls, a, b, c, d) || (n & 1)) return PBX_ERR_UNSUPPORTED;
    LmMapsP M;
    double *wb = ws, *wu = ws + (size_t)nl * (size_t)n;
    if (!lm_map(&M.a, a, n, nl, ls) || !lm_map(&M.b, b, n, nl, ls) || !lm_map(&M.c, c, n, nl, ls) ||
        !lm_map(&M.d, d, n, nl, ls) || !lm_map(&M.wb, wb, n, nl, n) || !lm_map(&M.wu, wu, n, nl, n))
        return PBX_ERR_UNSUPPORTED;
    periodic_lm_kernel<<<(unsigned)((nl + LM_LINES - 1) / LM_LINES), LM_LINES, sizeof(LmShared), s>>>(n, nl, ls, M,
                                                                                                  a, b, c);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
