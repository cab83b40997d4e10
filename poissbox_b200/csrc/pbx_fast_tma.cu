// pbx_fast_tma.cu -- TMA-pipelined persistent kernels of the FAST schedule (sm_100a).
//
// Same arithmetic as the generic kernels of pbx_fast_kernels.cu (bit-identical results: both are
// built from pbx_fast_common.cuh), different data movement:
//
//   * persistent CTAs (x pass: 2 per SM; y/z pass: 1 per SM with two compute groups), each
//     looping over tiles;
//   * thread 0 keeps the NEXT tile in flight with TMA (cp.async.bulk.tensor, completion on an
//     mbarrier) while all warps work on the current tile out of registers, so HBM
//     requests are outstanding all the time instead of only between a CTA's start and its first
//     barrier (the round-1a profile showed the generic kernels latency-bound: long-scoreboard
//     stalls, 24 % warp occupancy, 35-59 % DRAM utilisation);
//   * y / z pass: a tile is 16 x-columns x a whole line (x G lines), fetched as 3-D boxes
//     (16, n, G) / (16, G, n); threads pick their chunk and the 3-point stencil halos straight out
//     of the shared-memory tile; results go back with coalesced 64-byte row segments;
//   * x pass: a tile is 256 consecutive 16-point chunks (8 lines of 512) seen as a 2-D tensor
//     [chunks][16] and fetched with the 128-byte swizzle, which makes the one-chunk-per-lane
//     128-bit reads bank-conflict free; chunk states and halos travel by warp shuffles (a line is
//     at most one warp), and the two outputs leave through swizzled staging tiles and TMA stores.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "pbx_fast_common.cuh"
#include "pbx_fast_lineop.cuh"
#include "pbx_cg_dev.cuh"

namespace pbx {

using namespace fast;

namespace {

constexpr int XW = 8;
constexpr int TILE_DOUBLES = 4096;            // 32 KiB per field per tile
constexpr uint32_t TILE_BYTES = TILE_DOUBLES * 8;

using namespace ptx;   // pbx_ptx.cuh: mbarrier / TMA / fence wrappers

// ---------------------------------------------------------------------------------------------
// y / z pass.  One CTA per SM: a producer warp and two independent compute groups of 256 threads.
// A tile is 16 x-columns (128-byte rows, the granularity TMA moves at full rate; 64-byte rows ran
// the TMA unit out of row throughput) x a whole line x G lines; group g works on columns
// 8g .. 8g+7 with its own named barrier and exchange area, so the groups drift apart and one
// computes while the other waits at a barrier.
// ---------------------------------------------------------------------------------------------
constexpr int XWT = 16;                        // tile width in x
constexpr int NGRP = 2;                        // compute groups per CTA
constexpr int NTHR_YZ = NGRP * NT;             // no separate producer warp: see issue_tile()
constexpr int YZ_TILE_DOUBLES = 8192;          // 64 KiB per field per tile
constexpr uint32_t YZ_TILE_BYTES = YZ_TILE_DOUBLES * 8;

struct YZT {
    CompositeCoef M, D;
    int nx, n, T, G, ng;      // x extent, line length, chunks per line, lines per tile, lines total
    long long sl, sg;         // global strides along the line / between lines of a group
    int se, sgm;              // shared-memory strides (doubles) along the line / between lines
    int nbox, RB;             // boxes per field per tile, line points per box
    unsigned tbytes;          // bytes of one field's tile (64 KiB when the lines fill the CTA)
    int ntx, ntx8, ntiles;    // tiles along x (16 wide), 8-wide sub-tiles along x, total tiles
    int zdir;                 // 0: y pass (box = (16, RB, G)), 1: z pass (box = (16, G, RB))
    int rev;                  // 1: walk the tiles from the last to the first (L2 reuse, see lapl_fast)
    SegGeom seg;              // long lines: tile = (segment, x tile, line group), segment fastest
    int ngt;                  // line groups (tiles in the remaining direction)
};

struct TileId {
    int s, xt, gt;            // segment, 16-wide x tile, line group
};
__device__ __forceinline__ TileId tile_id(const YZT &p, int tile)
{
    const int s = tile % p.seg.nseg, rest = tile / p.seg.nseg;
    return {s, rest % p.ntx, rest / p.ntx};
}

struct YZShared {
    double tile[2][YZ_TILE_DOUBLES];
    double xchg[NGRP][Y_SLOTS * NT];
    uint64_t full, empty;
};

struct BarGroup {
    int id;
    __device__ __forceinline__ void operator()() const { named_bar_sync<NT>(id); }
};

// thread 0 doubles as the TMA producer: a 17th warp would put five warps on one SM sub-partition
// and cap the kernel at 96 registers per thread (16 K registers per sub-partition)
__device__ __forceinline__ void yz_issue_tile(YZShared &S, const YZT &p, const CUtensorMap *map0,
                                              const CUtensorMap *map1, int tile)
{
    if (p.rev) tile = p.ntiles - 1 - tile;
    const TileId id = tile_id(p, tile);
    const int x0 = id.xt * XWT, g0 = id.gt * p.G;
    // first line point of the tile: 0, or hlo chunks in front of the segment's interior; the boxes
    // of a segment tile wrap around the periodic line one by one (n is a multiple of RB)
    const int start = p.seg.nseg > 1 ? (id.s * p.seg.iseg - p.seg.hlo) * LC : 0;
    mbar_expect_tx(&S.full, 2 * p.tbytes);
    for (int b = 0; b < p.nbox; ++b) {
        int i0 = (start + b * p.RB) % p.n;
        if (i0 < 0) i0 += p.n;
        const int c1 = p.zdir ? g0 : i0, c2 = p.zdir ? i0 : g0;
        const int off = b * p.RB * p.se;
        tma_load_3d(&S.tile[0][off], map0, &S.full, x0, c1, c2);
        tma_load_3d(&S.tile[1][off], map1, &S.full, x0, c1, c2);
    }
}

// ROT (default; PBX_YZ_ROT=0 turns it off; 512^3: y pass 0.807 -> 0.786 ms, z pass 0.645 -> 0.626 ms): the tiles are fetched with the 128-byte swizzle and threads with an odd
// chunk index read the two halves of their chunk's second index bit in swapped order.  Without it the
// four chunk rows a warp reads per instruction lie 16 rows apart, i.e. in the same banks (the profile of
// round 1: 37-43 % of the passes' shared-memory wavefronts are bank conflicts); with it the two rows of
// a half-warp differ in bit 2 of (row & 7), the swizzle sends them to different halves of the 128-byte
// bank line, and 16 register swaps put the values back in order.  Whole lines in one CTA only (no
// segments, no slab), and for the z pass one line per tile row (G == 1: lines of 512 points).
// ANYT (default for such lengths; PBX_TMA_ANY_T=0 falls back to the generic kernels): line lengths whose chunk count does not divide 32 -- the lines that
// fit into a compute group leave threads without a chunk (`dead`); a compile-time switch, so that the
// measured kernels do not carry the test.
// FUSE (default, PBX_FUSE_TAIL=0 turns it off; z pass with the fused dot): the CTA that finishes last reduces the
// partial sums of p.out, all-reduces them over the peer boards and runs the CG's scalar step
// (cgdev::red_tail) -- compute and collective in ONE kernel, two launches fewer per iteration.
// DOT (z pass): the fused dot p . out -- a compile-time switch, so that the plain apply does not carry its registers.
template <bool ZPASS, bool SLAB, bool SEG, bool ROT, bool ANYT = false, bool FUSE = false, bool DOT = false>
__global__ void __launch_bounds__(NTHR_YZ, 1)
yz_tma_kernel(const __grid_constant__ YZT p, const __grid_constant__ ZOpen zo,
              const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
              double *__restrict__ out0, double *__restrict__ out1,
              const double *__restrict__ pv, double *__restrict__ partials,
              const __grid_constant__ RedTail tail)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    YZShared &S = *reinterpret_cast<YZShared *>(smraw);
    const int tid = threadIdx.x;
    if ((smem_u32(smraw) & (ROT ? 1023u : 127u)) != 0) __trap();   // TMA destinations: 128 bytes, swizzled 1 KiB

    if (tid == 0) {
        mbar_init(&S.full, 1);
        mbar_init(&S.empty, NTHR_YZ);
        fence_mbar_init();
        if ((int)blockIdx.x < p.ntiles) yz_issue_tile(S, p, &map0, &map1, blockIdx.x);
    }
    __syncthreads();

    // ===== compute groups =====
    const int grp = tid >> 8, lt = tid & (NT - 1);
    const int tx = lt & (XW - 1);
    const int t = (lt >> 3) % p.T;
    const int tz = lt / (XW * p.T);
    const BarGroup bar{1 + grp};
    constexpr bool segd = SEG;                               // tiles are segments of longer lines
    const int npts = SEG ? SEG_T * LC : p.n;                 // line points in a tile
    const bool dead = ANYT && tz >= p.G;                     // T does not divide 32: threads without a chunk
    Xchg xc{S.xchg[grp], lt, t, p.T, XW, (SLAB || segd) ? 1 : 0, dead ? 1 : 0};

    const int soff = tz * p.sgm + grp * XW + tx;            // + i * se
    int it = 0;
    for (int tile0 = blockIdx.x; tile0 < p.ntiles; tile0 += gridDim.x, ++it) {
        const int tile = p.rev ? p.ntiles - 1 - tile0 : tile0;
        TileId id{0, tile % p.ntx, tile / p.ntx};
        SegChunk sc{t, true};
        if (SEG) {
            id = tile_id(p, tile);
            sc = seg_chunk(p.seg, id.s, t);
        }
        const int xt8 = id.xt * NGRP + grp;                  // 8-wide sub-tile index along x
        const int gt = id.gt;
        const int x = xt8 * XW + tx;
        const int g = gt * p.G + tz;
        const bool live = (x < p.nx) && (g < p.ng) && sc.interior && !dead;
        const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;

        // slab: the message of the neighbour this thread's chunk touches (chunk 0: lower, chunk T-1: upper), issued
        // here so that the loads overlap the wait for the tile.  (Loading it one tile ahead was tried: no gain.)
        double m9[DIST_MSG];
        if (SLAB) slab_load_message(zo, t == 0, t == p.T - 1, live ? (long long)x + (long long)p.nx * g : 0, m9);
        mbar_wait(&S.full, (uint32_t)(it & 1));
        double a[LC], eb[LC + 6];
        if (ROT) {
            const int i0 = t * LC, odd = t & 1, flip = 4 * odd;
            const int x = grp * XW + tx;                       // column of the 16-wide tile
            const int qb = (ZPASS ? 0 : tz * npts) + i0;       // first tile row of my chunk (a multiple of 16)
            // load k fetches row qb + (k ^ flip); (row & 7) = (k & 7) ^ flip =: c ^ flip.  Eight base
            // offsets, one per c; the rest of the address is the immediate 16 k.
            int B[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                B[c] = (qb + ((c & 4) ? -flip : flip)) * 16 + ((((x >> 1) ^ c ^ flip) << 1) | (x & 1));
#pragma unroll
            for (int k = 0; k < LC; ++k) {
                a[k] = S.tile[0][B[k & 7] + 16 * k];
                eb[k + 3] = S.tile[1][B[k & 7] + 16 * k];
            }
#pragma unroll
            for (int k = 0; k < LC; ++k) {
                if (!(k & 4)) {     // the value fetched by load k belongs to index k ^ flip
                    const double a0 = a[k], a1 = a[k | 4], b0 = eb[k + 3], b1 = eb[(k | 4) + 3];
                    a[k] = odd ? a1 : a0;
                    a[k | 4] = odd ? a0 : a1;
                    eb[k + 3] = odd ? b1 : b0;
                    eb[(k | 4) + 3] = odd ? b0 : b1;
                }
            }
            const int q0 = (ZPASS ? 0 : tz * npts);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int il = i0 - 3 + k, ir = i0 + LC + k;
                if (il < 0) il += npts;
                if (ir >= npts) ir -= npts;
                const int ql = q0 + il, qr = q0 + ir;
                eb[k] = S.tile[1][ql * 16 + ((((x >> 1) ^ (ql & 7)) << 1) | (x & 1))];
                eb[LC + 3 + k] = S.tile[1][qr * 16 + ((((x >> 1) ^ (qr & 7)) << 1) | (x & 1))];
            }
        } else {
            const double *ta = S.tile[0] + soff, *tb = S.tile[1] + soff;
            const int i0 = t * LC;
#pragma unroll
            for (int k = 0; k < LC; ++k) a[k] = ta[(i0 + k) * p.se];
#pragma unroll
            for (int k = 0; k < LC; ++k) eb[k + 3] = tb[(i0 + k) * p.se];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int il = i0 - 3 + k, ir = i0 + LC + k;
                const bool lo = il < 0, hi = ir >= npts;
                if (lo) il += npts;
                if (hi) ir -= npts;
                const double vl = tb[il * p.se], vr = tb[ir * p.se];
                const bool cut = SLAB || segd;    // open line: nothing beyond the slab / segment
                eb[k] = (cut && lo) ? 0.0 : vl;
                eb[LC + 3 + k] = (cut && hi) ? 0.0 : vr;
            }
        }
        // The release of the tile buffers.  fence.proxy.async FIRST: the mbarrier arrival does not wait for this
        // thread's outstanding shared-memory loads (SASS: 38 LDS, then SYNCS.ARRIVE with no scoreboard wait in
        // between), and the TMA writes it releases belong to the async proxy, which the arrival's release semantics
        // do not order against generic-proxy reads.  Without the fence the next tile's boxes could land under loads
        // still queued behind bank conflicts: round 2 saw exactly that with the segment tiles (one 64-row box of a
        // (tile, group) read as the NEXT tile's data in 3 of 40 applies, tools/seg_defect_probe2.py; 0 of 40 with the
        // fence, 0 of 40 with an arrival made data-dependent on every value read).
        // Cost at 512^3: 0.4 % of the apply (profiles/r2_fence_cost_and_seg_tma.log).
        fence_proxy_async();
        mbar_arrive(&S.empty);   // this thread no longer needs the tile buffers
        if (tid == 0 && tile0 + (int)gridDim.x < p.ntiles) {
            mbar_wait(&S.empty, (uint32_t)(it & 1));       // ... and neither does anybody else
            yz_issue_tile(S, p, &map0, &map1, tile0 + gridDim.x);
        }

        if (!ZPASS) {
            double c[LC], d[LC];
            ypass_body(p.M, p.D, xc, a, eb, c, d, bar);
            if (live) {
#pragma unroll
                for (int k = 0; k < LC; ++k) {
                    out0[base + k * p.sl] = c[k];
                    out1[base + k * p.sl] = d[k];
                }
            }
        } else {
            double o[LC], pk[LC];
            // fused dot: the rows of p are loaded just before the pass's last barrier (zpass_body's hook) -- after it
            // they would cost a DRAM round trip per tile with every warp of the group waiting
            auto load_p = [&]() {
                if (DOT && live) {
#pragma unroll
                    for (int k = 0; k < LC; ++k) pk[k] = __ldg(pv + base + k * p.sl);
                }
            };
            if (SLAB) {
                zpass_body_slab(p.M, p.D, zo, xc, m9, m9, a, eb, o, bar, load_p);
            } else {
                zpass_body(p.M, p.D, xc, a, eb, o, bar, load_p);
            }
            double dot = 0.0;
            if (live) {
                if (DOT) {
#pragma unroll
                    for (int k = 0; k < LC; ++k) {
                        dot = fma(pk[k], o[k], dot);
                        out0[base + k * p.sl] = o[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < LC; ++k) out0[base + k * p.sl] = o[k];
                }
            }
            if (DOT) {
                // one partial per 8-wide sub-tile, numbered as the generic kernel numbers its CTAs;
                // the warp sums go through the last exchange slot, which the z pass never uses
                double tot = block_sum_warps(dot, S.xchg[grp] + (Y_SLOTS - 1) * NT, lt, NT, bar);
                if (lt == 0 && xt8 < p.ntx8) partials[(id.s * p.ngt + gt) * p.ntx8 + xt8] = tot;
            }
        }
    }
    if (FUSE) cgdev::red_tail<NTHR_YZ>(tail);
}

// ---------------------------------------------------------------------------------------------
// x pass
// ---------------------------------------------------------------------------------------------
struct XT {
    CompositeCoef M, D;
    int T;                    // chunks per line (power of two; ANYT: any number up to 256)
    int ntiles;               // tiles of 256 chunks (ANYT: of `rows` chunks)
    int rev;                  // 1: walk the tiles from the last to the first
    int rows;                 // ANYT: chunks per tile = the whole lines that fit into 256
};

struct XShared {
    double tin[TILE_DOUBLES];
    double sta[TILE_DOUBLES];   // WIDE: the chunk-state exchange area (X_SLOTS * NT doubles) lives in
    double stb[TILE_DOUBLES];   // sta..stb while a tile is computed, before the results are staged
    uint64_t full, empty;
};
static_assert(X_SLOTS * NT <= 2 * TILE_DOUBLES, "exchange area must fit the two staging tiles");

// swizzled address of 16-byte piece j of row q (128-byte rows, TMA SWIZZLE_128B)
__device__ __forceinline__ int swz(int q, int j) { return q * 16 + ((j ^ (q & 7)) << 1); }

__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

// shuffle counterpart of fast::lookback: states live in registers of the lanes of one line
__device__ __forceinline__ void lookback_shfl(const CompositeCoef &c, double ey, double ez, int lane,
                                              int T, int dir, double &Y, double &Z)
{
    const int seg = lane & ~(T - 1), t = lane & (T - 1);
    int src = seg | ((t + dir) & (T - 1));
    Y = shfl_d(ey, src);
    Z = shfl_d(ez, src);
#pragma unroll
    for (int m = 2; m <= MAXLOOK; ++m) {
        if (m <= c.nlook) {     // uniform across the warp
            src = seg | ((t + dir * m) & (T - 1));
            double y2 = shfl_d(ey, src), z2 = shfl_d(ez, src);
            double pp = c.look[m - 1];
            Y = fma(pp, y2, Y);
            Z = fma(pp, fma((double)(LC * (m - 1)), y2, z2), Z);
        }
    }
}

__device__ __forceinline__ void solve_shfl(const CompositeCoef &c, double (&v)[LC], int lane, int T)
{
    double ey, ez, Y, Z;
    fwd_local(c.r, v, ey, ez);
    lookback_shfl(c, ey, ez, lane, T, -1, Y, Z);
    fwd_fix(c, v, Y, Z);
    bwd_local(c.r, v, ey, ez);
    lookback_shfl(c, ey, ez, lane, T, +1, Y, Z);
    bwd_fix(c, v, Y, Z);
}

// WIDE = false: a line is at most one warp (T <= 32 chunks); states and halos by shuffle.
// WIDE = true:  T = 64, 128 or 256 chunks (lines of 1024 - 4096 points): a line spans several warps
//               of the CTA, and the chunk states and solved halos go through shared memory
//               (xpass_body, the arithmetic of the generic x kernel) -- same TMA data movement.
// ANYT (default for such lengths; with WIDE): any number of chunks per line up to 256 -- a tile holds
//               the whole lines that fit into 256 chunks, the remaining threads idle.
template <bool WIDE, bool ANYT = false>
__global__ void __launch_bounds__(NT, 2)
x_tma_kernel(const __grid_constant__ XT p, const __grid_constant__ CUtensorMap mapF,
             const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB)
{
    static_assert(WIDE || !ANYT, "ANYT runs on the shared-memory exchange of the WIDE kernel");
    const int rows = ANYT ? p.rows : NT;
    const uint32_t tile_bytes = ANYT ? (uint32_t)p.rows * 128u : TILE_BYTES;
    extern __shared__ __align__(1024) unsigned char smraw[];
    XShared &S = *reinterpret_cast<XShared *>(smraw);
    const int tid = threadIdx.x;
    if ((smem_u32(smraw) & 1023u) != 0) __trap();  // the 128-byte swizzle pattern repeats every 1 KiB
    if (tid == 0) {
        mbar_init(&S.full, 1);
        mbar_init(&S.empty, NT);
        fence_mbar_init();
        if ((int)blockIdx.x < p.ntiles) {
            mbar_expect_tx(&S.full, tile_bytes);
            tma_load_2d(S.tin, &mapF, &S.full, 0, (p.rev ? p.ntiles - 1 - (int)blockIdx.x : (int)blockIdx.x) * rows);
        }
    }
    __syncthreads();

    const int lane = tid & 31;
    const int T = p.T;
    const int seg = lane & ~(T - 1), t = lane & (T - 1);
    // chunk of the line and first row of the line (WIDE), rows of the previous / next chunk of the line
    const int tw = ANYT ? tid % T : (tid & (T - 1)), lb = tid - tw;
    const bool dead = ANYT && tid >= rows;
    const int ql = ANYT ? lb + (tw == 0 ? T - 1 : tw - 1)
                        : WIDE ? ((tid & ~(T - 1)) | ((tid - 1) & (T - 1))) : ((tid & ~31) | seg | ((t - 1) & (T - 1)));
    const int qr = ANYT ? lb + (tw == T - 1 ? 0 : tw + 1)
                        : WIDE ? ((tid & ~(T - 1)) | ((tid + 1) & (T - 1))) : ((tid & ~31) | seg | ((t + 1) & (T - 1)));
    int it = 0;
    for (int tile0 = blockIdx.x; tile0 < p.ntiles; tile0 += gridDim.x, ++it) {
        const int tile = p.rev ? p.ntiles - 1 - tile0 : tile0;
        mbar_wait(&S.full, (uint32_t)(it & 1));
        double ef[LC + 6];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double2 v2 = *reinterpret_cast<const double2 *>(S.tin + swz(tid, j));
            ef[3 + 2 * j] = v2.x;
            ef[4 + 2 * j] = v2.y;
        }
        {
            // (an idle ANYT thread's neighbours lie inside the tile buffer too: lb + T - 1 < 256 + T <= 512 rows
            //  of tin + sta, never beyond the shared-memory block)
            double2 l6 = *reinterpret_cast<const double2 *>(S.tin + swz(ql, 6));
            double2 l7 = *reinterpret_cast<const double2 *>(S.tin + swz(ql, 7));
            double2 r0 = *reinterpret_cast<const double2 *>(S.tin + swz(qr, 0));
            double2 r1 = *reinterpret_cast<const double2 *>(S.tin + swz(qr, 1));
            ef[0] = l6.y;
            ef[1] = l7.x;
            ef[2] = l7.y;
            ef[LC + 3] = r0.x;
            ef[LC + 4] = r0.y;
            ef[LC + 5] = r1.x;
        }
        fence_proxy_async();   // generic-proxy reads before the async-proxy refill, see yz_tma_kernel
        mbar_arrive(&S.empty);
        if (tid == 0 && tile0 + (int)gridDim.x < p.ntiles) {
            const int nxt = tile0 + (int)gridDim.x;
            mbar_wait(&S.empty, (uint32_t)(it & 1));
            mbar_expect_tx(&S.full, tile_bytes);
            tma_load_2d(S.tin, &mapF, &S.full, 0, (p.rev ? p.ntiles - 1 - nxt : nxt) * rows);
        }

        double va[LC], vb[LC];
        if (WIDE) {
            // the exchange area aliases the staging tiles: the previous tile's TMA stores must have
            // read them
            if (tid == 0) tma_wait_read0();
            BarCompute()();
            const Xchg xc{S.sta, tid, tw, T, 1, 0, dead ? 1 : 0};
            xpass_body(p.M, p.D, xc, ef, va, vb, BarCompute());
        } else {
        stencil<true>(p.D, ef, va);
#pragma unroll
        for (int k = 0; k < LC; ++k) vb[k] = ef[k + 3];
        solve_shfl(p.D, va, lane, T);
        solve_shfl(p.M, vb, lane, T);
        {
            // S_M on the solved values: halos of the neighbouring chunks by shuffle
            double e[LC + 6];
            const int sl = seg | ((t - 1) & (T - 1)), sr = seg | ((t + 1) & (T - 1));
            e[0] = shfl_d(vb[LC - 3], sl);
            e[1] = shfl_d(vb[LC - 2], sl);
            e[2] = shfl_d(vb[LC - 1], sl);
            e[LC + 3] = shfl_d(vb[0], sr);
            e[LC + 4] = shfl_d(vb[1], sr);
            e[LC + 5] = shfl_d(vb[2], sr);
#pragma unroll
            for (int k = 0; k < LC; ++k) e[k + 3] = vb[k];
            stencil<false>(p.M, e, vb);
        }
        }

        // staging tiles: wait until the previous tile's TMA stores have read them (WIDE: until
        // everybody has read the exchange area)
        if (!WIDE && tid == 0) tma_wait_read0();
        BarCompute()();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<double2 *>(S.sta + swz(tid, j)) = make_double2(va[2 * j], va[2 * j + 1]);
            *reinterpret_cast<double2 *>(S.stb + swz(tid, j)) = make_double2(vb[2 * j], vb[2 * j + 1]);
        }
        fence_proxy_async();
        BarCompute()();
        if (tid == 0) {
            tma_store_2d(&mapA, S.sta, 0, tile * rows);
            tma_store_2d(&mapB, S.stb, 0, tile * rows);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all0();
}

// ---------------------------------------------------------------------------------------------
// ONE compact line operator per launch (grad / div / interp of the FAST schedule, pbx_fast_lineop.cu)
// with the data movement of the Laplacian passes.  A single input field leaves room for TWO tile
// stages, so the tile after next is in flight while the current one is computed.  Arithmetic from
// pbx_fast_lineop.cuh: same bits as the generic line-operator kernels.  Default since round 2
// (PBX_LINEOP_TMA=0: the generic kernels).
// ---------------------------------------------------------------------------------------------
struct LYZShared {
    double tile[2][YZ_TILE_DOUBLES];
    double xchg[NGRP][4 * NT];
    uint64_t full[2], empty[2];
};

__device__ __forceinline__ void lyz_issue_tile(LYZShared &S, int stage, const YZT &p, const CUtensorMap *map,
                                               int tile)
{
    const TileId id = tile_id(p, tile);
    const int x0 = id.xt * XWT, g0 = id.gt * p.G;
    // a segment of a longer line starts hlo chunks in front of its interior; its boxes wrap around the
    // periodic line one by one (yz_issue_tile)
    const int start = p.seg.nseg > 1 ? (id.s * p.seg.iseg - p.seg.hlo) * LC : 0;
    mbar_expect_tx(&S.full[stage], p.tbytes);
    for (int b = 0; b < p.nbox; ++b) {
        int i0 = (start + b * p.RB) % p.n;
        if (i0 < 0) i0 += p.n;
        const int c1 = p.zdir ? g0 : i0, c2 = p.zdir ? i0 : g0;
        tma_load_3d(&S.tile[stage][b * p.RB * p.se], map, &S.full[stage], x0, c1, c2);
    }
}

// DUAL: two operators on the same input per tile (grad's interpolation and derivative of one field along
// z, resp. y: the input is read once, 24 instead of 32 B/point for the pair); second result to out2.
template <bool ADD, bool SEG, bool DUAL = false>
__global__ void __launch_bounds__(NTHR_YZ, 1)
lineop_yz_tma_kernel(const __grid_constant__ YZT p, const __grid_constant__ lineop::LineOp op,
                     const __grid_constant__ CUtensorMap map, const double *__restrict__ addend,
                     double *__restrict__ out, const __grid_constant__ lineop::LineOp op2,
                     double *__restrict__ out2)
{
    static_assert(!(ADD && DUAL), "the dual kernel has no addend");
    extern __shared__ __align__(1024) unsigned char smraw[];
    LYZShared &S = *reinterpret_cast<LYZShared *>(smraw);
    const int tid = threadIdx.x;
    if ((smem_u32(smraw) & 127u) != 0) __trap();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], NTHR_YZ);
        }
        fence_mbar_init();
        for (int s = 0; s < 2; ++s)
            if ((int)blockIdx.x + s * (int)gridDim.x < p.ntiles)
                lyz_issue_tile(S, s, p, &map, blockIdx.x + s * gridDim.x);
    }
    __syncthreads();
    const int grp = tid >> 8, lt = tid & (NT - 1);
    const int tx = lt & (XW - 1);
    const int t = (lt >> 3) % p.T;
    const int tz = lt / (XW * p.T);
    const BarGroup bar{1 + grp};
    const bool dead = tz >= p.G;
    const Xchg xc{S.xchg[grp], lt, t, p.T, XW, SEG ? 1 : 0, dead ? 1 : 0};
    const int soff = (dead ? 0 : tz * p.sgm) + grp * XW + tx;
    const int npts = SEG ? SEG_T * LC : p.n;     // line points in a tile
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        TileId id{0, tile % p.ntx, tile / p.ntx};
        SegChunk sc{t, true};
        if (SEG) {
            id = tile_id(p, tile);
            sc = seg_chunk(p.seg, id.s, t);
        }
        const int xt8 = id.xt * NGRP + grp, gt = id.gt;
        const int x = xt8 * XW + tx, g = gt * p.G + tz;
        const bool live = (x < p.nx) && (g < p.ng) && !dead && sc.interior;
        const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;
        mbar_wait(&S.full[st], par);
        double e[LC + 6], v[LC];
        {
            const double *tb = S.tile[st] + soff;
            const int i0 = t * LC;
#pragma unroll
            for (int k = 0; k < LC; ++k) e[k + 3] = tb[(i0 + k) * p.se];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int il = i0 - 3 + k, ir = i0 + LC + k;
                const bool lo = il < 0, hi = ir >= npts;
                if (lo) il += npts;
                if (hi) ir -= npts;
                const double vl = tb[il * p.se], vr = tb[ir * p.se];
                e[k] = (SEG && lo) ? 0.0 : vl;            // a segment is an open line
                e[LC + 3 + k] = (SEG && hi) ? 0.0 : vr;
            }
        }
        fence_proxy_async();   // generic-proxy reads before the async-proxy refill, see yz_tma_kernel
        mbar_arrive(&S.empty[st]);
        if (tid == 0 && tile + 2 * (int)gridDim.x < p.ntiles) {
            mbar_wait(&S.empty[st], par);
            lyz_issue_tile(S, st, p, &map, tile + 2 * gridDim.x);
        }
        lineop::stencil4(op, e, v);
        lineop::solve1_chunk(op.cc, xc, 0, v, bar);
        if (live) {
            if (ADD) {
#pragma unroll
                for (int k = 0; k < LC; ++k) out[base + k * p.sl] = v[k] + __ldg(addend + base + k * p.sl);
            } else {
#pragma unroll
                for (int k = 0; k < LC; ++k) out[base + k * p.sl] = v[k];
            }
        }
        if (DUAL) {
            lineop::stencil4(op2, e, v);
            lineop::solve1_chunk(op2.cc, xc, 2, v, bar);
            if (live) {
#pragma unroll
                for (int k = 0; k < LC; ++k) out2[base + k * p.sl] = v[k];
            }
        }
    }
}

// out = opA(inA) + opB(inB) along y or z: div's sums (src/compact_schemes.f90:249, 251) without the
// intermediate field -- the first summand never leaves the registers (24 instead of 40 B/point for the
// pair).  Two input tiles, one stage (the tile layout of the Laplacian's y / z pass).  The first
// operator's chunk is solved before the second field is taken out of the tile, which bounds the
// registers; the sum is formed as the generic kernels form it (second operator + first).
struct LYZ2Shared {
    double tile[2][YZ_TILE_DOUBLES];
    double xchg[NGRP][4 * NT];
    uint64_t full, empty;
};

template <bool SEG>
__global__ void __launch_bounds__(NTHR_YZ, 1)
lineop_yz_tma_sum_kernel(const __grid_constant__ YZT p, const __grid_constant__ lineop::LineOp opA,
                         const __grid_constant__ lineop::LineOp opB, const __grid_constant__ CUtensorMap mapA,
                         const __grid_constant__ CUtensorMap mapB, double *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    LYZ2Shared &S = *reinterpret_cast<LYZ2Shared *>(smraw);
    const int tid = threadIdx.x;
    if ((smem_u32(smraw) & 127u) != 0) __trap();
    auto issue = [&](int tile) {
        const TileId id = tile_id(p, tile);
        const int x0 = id.xt * XWT, g0 = id.gt * p.G;
        const int start = p.seg.nseg > 1 ? (id.s * p.seg.iseg - p.seg.hlo) * LC : 0;
        mbar_expect_tx(&S.full, 2 * p.tbytes);
        for (int b = 0; b < p.nbox; ++b) {
            int i0 = (start + b * p.RB) % p.n;
            if (i0 < 0) i0 += p.n;
            const int c1 = p.zdir ? g0 : i0, c2 = p.zdir ? i0 : g0;
            tma_load_3d(&S.tile[0][b * p.RB * p.se], &mapA, &S.full, x0, c1, c2);
            tma_load_3d(&S.tile[1][b * p.RB * p.se], &mapB, &S.full, x0, c1, c2);
        }
    };
    if (tid == 0) {
        mbar_init(&S.full, 1);
        mbar_init(&S.empty, NTHR_YZ);
        fence_mbar_init();
        if ((int)blockIdx.x < p.ntiles) issue(blockIdx.x);
    }
    __syncthreads();
    const int grp = tid >> 8, lt = tid & (NT - 1);
    const int tx = lt & (XW - 1);
    const int t = (lt >> 3) % p.T;
    const int tz = lt / (XW * p.T);
    const BarGroup bar{1 + grp};
    const bool dead = tz >= p.G;
    const Xchg xc{S.xchg[grp], lt, t, p.T, XW, SEG ? 1 : 0, dead ? 1 : 0};
    const int soff = (dead ? 0 : tz * p.sgm) + grp * XW + tx;
    const int npts = SEG ? SEG_T * LC : p.n;
    auto take = [&](const double *tb, double (&e)[LC + 6]) {
        const int i0 = t * LC;
#pragma unroll
        for (int k = 0; k < LC; ++k) e[k + 3] = tb[(i0 + k) * p.se];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int il = i0 - 3 + k, ir = i0 + LC + k;
            const bool lo = il < 0, hi = ir >= npts;
            if (lo) il += npts;
            if (hi) ir -= npts;
            const double vl = tb[il * p.se], vr = tb[ir * p.se];
            e[k] = (SEG && lo) ? 0.0 : vl;
            e[LC + 3 + k] = (SEG && hi) ? 0.0 : vr;
        }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        TileId id{0, tile % p.ntx, tile / p.ntx};
        SegChunk sc{t, true};
        if (SEG) {
            id = tile_id(p, tile);
            sc = seg_chunk(p.seg, id.s, t);
        }
        const int xt8 = id.xt * NGRP + grp, gt = id.gt;
        const int x = xt8 * XW + tx, g = gt * p.G + tz;
        const bool live = (x < p.nx) && (g < p.ng) && !dead && sc.interior;
        const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;
        mbar_wait(&S.full, (uint32_t)(it & 1));
        double e[LC + 6], va[LC], vb[LC];
        take(S.tile[0] + soff, e);
        lineop::stencil4(opA, e, va);
        lineop::solve1_chunk(opA.cc, xc, 0, va, bar);
        take(S.tile[1] + soff, e);
        fence_proxy_async();   // generic-proxy reads before the async-proxy refill, see yz_tma_kernel
        mbar_arrive(&S.empty);
        if (tid == 0 && tile + (int)gridDim.x < p.ntiles) {
            mbar_wait(&S.empty, (uint32_t)(it & 1));
            issue(tile + gridDim.x);
        }
        lineop::stencil4(opB, e, vb);
        lineop::solve1_chunk(opB.cc, xc, 2, vb, bar);
        if (live) {
#pragma unroll
            for (int k = 0; k < LC; ++k) out[base + k * p.sl] = vb[k] + va[k];
        }
    }
}

struct LXShared {
    double tin[2][TILE_DOUBLES];
    double sta[TILE_DOUBLES];
    double xchg[2 * NT];          // WIDE: chunk states of lines that span several warps
    uint64_t full[2], empty[2];
};

// single-pole look-back by shuffle: the order of lineop::lookback1
__device__ __forceinline__ double lookback1_shfl(const CompositeCoef &c, double e, int lane, int T, int dir)
{
    const int seg = lane & ~(T - 1), t = lane & (T - 1);
    double S = shfl_d(e, seg | ((t + dir) & (T - 1)));
#pragma unroll
    for (int m = 2; m <= MAXLOOK; ++m)
        if (m <= c.nlook) S = fma(c.look[m - 1], shfl_d(e, seg | ((t + dir * m) & (T - 1))), S);
    return S;
}

struct LXT {
    lineop::LineOp op;
    int T, ntiles;
};

// WIDE: lines of 1024 - 4096 points (64 - 256 chunks) span several warps; their chunk states go through
// shared memory (lineop::solve1_chunk) instead of shuffles.
template <bool WIDE>
__global__ void __launch_bounds__(NT, 2)
lineop_x_tma_kernel(const __grid_constant__ LXT p, const __grid_constant__ CUtensorMap mapIn,
                    const __grid_constant__ CUtensorMap mapOut)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    LXShared &S = *reinterpret_cast<LXShared *>(smraw);
    const int tid = threadIdx.x;
    if ((smem_u32(smraw) & 1023u) != 0) __trap();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], NT);
        }
        fence_mbar_init();
        for (int s = 0; s < 2; ++s) {
            const int tl = (int)blockIdx.x + s * (int)gridDim.x;
            if (tl < p.ntiles) {
                mbar_expect_tx(&S.full[s], TILE_BYTES);
                tma_load_2d(S.tin[s], &mapIn, &S.full[s], 0, tl * NT);
            }
        }
    }
    __syncthreads();
    const int lane = tid & 31, T = p.T;
    const int seg = lane & ~(T - 1), t = lane & (T - 1);
    const int ql = WIDE ? ((tid & ~(T - 1)) | ((tid - 1) & (T - 1))) : ((tid & ~31) | seg | ((t - 1) & (T - 1)));
    const int qr = WIDE ? ((tid & ~(T - 1)) | ((tid + 1) & (T - 1))) : ((tid & ~31) | seg | ((t + 1) & (T - 1)));
    const CompositeCoef &c = p.op.cc;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t par = (uint32_t)((it >> 1) & 1);
        mbar_wait(&S.full[st], par);
        double e[LC + 6], v[LC];
        const double *tin = S.tin[st];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 v2 = *reinterpret_cast<const double2 *>(tin + swz(tid, j));
            e[3 + 2 * j] = v2.x;
            e[4 + 2 * j] = v2.y;
        }
        {
            const double2 l6 = *reinterpret_cast<const double2 *>(tin + swz(ql, 6));
            const double2 l7 = *reinterpret_cast<const double2 *>(tin + swz(ql, 7));
            const double2 r0 = *reinterpret_cast<const double2 *>(tin + swz(qr, 0));
            const double2 r1 = *reinterpret_cast<const double2 *>(tin + swz(qr, 1));
            e[0] = l6.y;
            e[1] = l7.x;
            e[2] = l7.y;
            e[LC + 3] = r0.x;
            e[LC + 4] = r0.y;
            e[LC + 5] = r1.x;
        }
        fence_proxy_async();   // generic-proxy reads before the async-proxy refill, see yz_tma_kernel
        mbar_arrive(&S.empty[st]);
        if (tid == 0 && tile + 2 * (int)gridDim.x < p.ntiles) {
            mbar_wait(&S.empty[st], par);
            mbar_expect_tx(&S.full[st], TILE_BYTES);
            tma_load_2d(S.tin[st], &mapIn, &S.full[st], 0, (tile + 2 * (int)gridDim.x) * NT);
        }
        lineop::stencil4(p.op, e, v);
        if (WIDE) {
            // (two barriers per tile order the exchange slots between consecutive tiles)
            const Xchg xc{S.xchg, tid, tid & (T - 1), T, 1};
            lineop::solve1_chunk(c, xc, 0, v, BarCompute());
        } else {
        // lineop::solve1_chunk with the chunk states travelling by shuffle
        double y = 0.0;
#pragma unroll
        for (int k = 0; k < LC; ++k) {
            y = fma(c.r, y, v[k]);
            v[k] = y;
        }
        const double Sc = lookback1_shfl(c, y, lane, T, -1);
        double w = 0.0;
#pragma unroll
        for (int k = LC - 1; k >= 0; --k) {
            const double yk = fma(c.pw[k], Sc, v[k]);
            w = fma(c.r, w, yk);
            v[k] = w;
        }
        const double Wc = lookback1_shfl(c, w, lane, T, +1);
#pragma unroll
        for (int k = 0; k < LC; ++k) v[k] = fma(c.pw[LC - 1 - k], Wc, v[k]);
        }
        // staging tile: the previous tile's TMA store must have read it
        if (tid == 0) tma_wait_read0();
        BarCompute()();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<double2 *>(S.sta + swz(tid, j)) = make_double2(v[2 * j], v[2 * j + 1]);
        fence_proxy_async();
        BarCompute()();
        if (tid == 0) {
            tma_store_2d(&mapOut, S.sta, 0, tile * NT);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all0();
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// 3-D map of a brick for the y / z pass tiles
bool make_map_yz(CUtensorMap *m, const double *base, const Brick &g, const YZT &p, bool swizzle = false)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)g.nz};
    cuuint64_t strides[2] = {(cuuint64_t)g.nx * 8, (cuuint64_t)g.nx * g.ny * 8};
    cuuint32_t box[3] = {(cuuint32_t)XWT, (cuuint32_t)(p.zdir ? p.G : p.RB),
                         (cuuint32_t)(p.zdir ? p.RB : p.G)};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box,
              es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D map [chunks][16] with the 128-byte swizzle for the x pass
bool make_map_x(CUtensorMap *m, const double *base, size_t nchunks, int rows = NT)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {16, (cuuint64_t)nchunks};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {16, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box,
              es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

bool yz_geometry_tma(const Brick &g, int dir, YZT *p)
{
    const int n = dir == 1 ? g.ny : g.nz;
    if (n % LC || n < LC || (g.nx & 1)) return false;
    p->nx = g.nx;
    p->n = n;
    p->seg = seg_geometry(n / LC);
    p->T = p->seg.T;
    // Lines of more than 512 points run as segment tiles (eight wrapped 64-point boxes per field).  Round 2 found these
    // tiles returning a wrong (tile, group) in 10-30 % of the applies on large bricks; cause and fix are at the release
    // of the tile buffers in yz_tma_kernel (proxy fence).  Since the fix: 0 of 250 applies differ from the generic
    // kernels (profiles/r2_seg_defect_fix_confirm.log), and the TMA kernels are 12-17 % faster on such bricks
    // (profiles/r2_fence_cost_and_seg_tma.log).  PBX_TMA_SEG=0 sends segmented lines to the generic kernels.
    if (p->seg.nseg > 1 && !env_switch("PBX_TMA_SEG", true)) return false;
    // T divides 32 -- or the lines that fit leave some threads of the group without a chunk (any multiple
    // of 16 up to 512 points; 384^3: 51.6 instead of 37.4 GDoF/s on the generic kernels; PBX_TMA_ANY_T=0
    // turns it off)
    if (NT % (XW * p->T)) {
        if (!env_switch("PBX_TMA_ANY_T", true) || p->seg.nseg > 1) return false;
    }
    p->G = NT / (XW * p->T);
    p->ng = dir == 1 ? g.nz : g.ny;
    p->sl = dir == 1 ? (long long)g.nx : (long long)g.nx * g.ny;
    p->sg = dir == 1 ? (long long)g.nx * g.ny : (long long)g.nx;
    p->zdir = dir == 2;
    const int npts = p->T * LC;                  // line points in a tile
    if (p->seg.nseg > 1) {
        // segment tiles start at multiples of 4 chunks: 64-point boxes never straddle the line end
        if (n % 64) return false;
        p->RB = 64;
        p->nbox = npts / p->RB;
    } else {
        p->nbox = n > 256 ? 2 : 1;
        p->RB = n / p->nbox;
    }
    if (p->nbox > 1 && p->G != 1) return false;
    p->tbytes = (unsigned)(XWT * npts * p->G * 8);
    if (p->tbytes > YZ_TILE_BYTES) return false;
    if (p->zdir) {          // smem layout [i][g][16]
        p->se = p->G * XWT;
        p->sgm = XWT;
    } else {                // smem layout [g][i][16]
        p->se = XWT;
        p->sgm = npts * XWT;
    }
    p->ntx = (g.nx + XWT - 1) / XWT;
    p->ntx8 = (g.nx + XW - 1) / XW;
    p->ngt = (p->ng + p->G - 1) / p->G;
    const long long nt = (long long)p->ntx * p->ngt * p->seg.nseg;
    if (nt > 0x7fffffffLL) return false;
    p->ntiles = (int)nt;
    return true;
}

}  // namespace

bool fast_tma_available() { return encode_fn() != nullptr; }


// 2-D fp64 tensor map {dim0 (contiguous), dim1} with row stride `stride1_bytes` (a multiple of 16)
bool tma_make_map_2d(CUtensorMap_st *m, const double *base, unsigned long long dim0, unsigned long long dim1,
                     unsigned long long stride1_bytes, unsigned box0, unsigned box1, bool swizzle128)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns PBX_ERR_UNSUPPORTED when the shape does not fit the TMA kernels (caller falls back)
int fast_xpass_tma(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *f, double *A,
                   double *B, int rev, long long *launches)
{
    const int T = g.nx / LC;
    if (g.nx % LC || T > NT || T < 1 || !encode_fn()) return PBX_ERR_UNSUPPORTED;
    const bool anyT = (T & (T - 1)) != 0;   // chunk counts that are not a power of two
    if (anyT && !env_switch("PBX_TMA_ANY_T", true)) return PBX_ERR_UNSUPPORTED;
    const bool wide = T > 32;
    const size_t nchunks = g.N() / LC;
    if (nchunks > 0x7fffffffull) return PBX_ERR_UNSUPPORTED;
    XT p;
    p.M = fc.M;
    p.D = fc.D[0];
    p.T = T;
    p.rows = anyT ? (NT / T) * T : NT;
    p.ntiles = (int)((nchunks + p.rows - 1) / p.rows);
    p.rev = rev;
    CUtensorMap mf, ma, mb;
    if (!make_map_x(&mf, f, nchunks, p.rows) || !make_map_x(&ma, A, nchunks, p.rows) ||
        !make_map_x(&mb, B, nchunks, p.rows))
        return PBX_ERR_UNSUPPORTED;
    const size_t smem = sizeof(XShared);
    static std::atomic<bool> attr_set[64];   // per device: the attribute belongs to the context
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBX_CUDA(cudaFuncSetAttribute(x_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        PBX_CUDA(cudaFuncSetAttribute(x_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        PBX_CUDA(cudaFuncSetAttribute(x_tma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set[dev_ & 63] = true;
    }
    int grid = 2 * sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    if (anyT)
        x_tma_kernel<true, true><<<grid, NT, smem, s>>>(p, mf, ma, mb);
    else if (wide)
        x_tma_kernel<true><<<grid, NT, smem, s>>>(p, mf, ma, mb);
    else
        x_tma_kernel<false><<<grid, NT, smem, s>>>(p, mf, ma, mb);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// one instantiation of the y / z kernel: its shared-memory attribute (once per device) and its launch
template <bool ZPASS, bool SLAB, bool SEG, bool ROT, bool ANYT, bool FUSE, bool DOT>
static int launch_yz(cudaStream_t s, int grid, size_t smem, const YZT &p, const ZOpen &zo, const CUtensorMap &m0,
                     const CUtensorMap &m1, double *out0, double *out1, const double *pv, double *partials,
                     const RedTail &tail)
{
    static std::atomic<bool> attr_set[64];   // per device: the attribute belongs to the context
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBX_CUDA(cudaFuncSetAttribute(yz_tma_kernel<ZPASS, SLAB, SEG, ROT, ANYT, FUSE, DOT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev_ & 63] = true;
    }
    yz_tma_kernel<ZPASS, SLAB, SEG, ROT, ANYT, FUSE, DOT><<<grid, NTHR_YZ, smem, s>>>(p, zo, m0, m1, out0, out1, pv, partials, tail);
    return PBX_OK;
}

int fast_yzpass_tma(cudaStream_t s, const Brick &g, const FastCoefs &fc, int dir, const double *in0,
                    const double *in1, double *out0, double *out1, const double *pvec,
                    double *partials, const ZOpen &zo, int rev, long long *launches, const RedTail *tail,
                    bool *tail_used)
{
    if (tail_used) *tail_used = false;
    YZT p;
    if (!encode_fn() || !yz_geometry_tma(g, dir, &p)) return PBX_ERR_UNSUPPORTED;
    if (zo.open && p.seg.nseg > 1) return PBX_ERR_UNSUPPORTED;   // the generic launcher reports it
    const bool anyT = NT % (XW * p.T) != 0;
    if (zo.open && anyT) return PBX_ERR_UNSUPPORTED;             // the slab look-back indexes directly
    if (zo.open && p.T < 2) return PBX_ERR_UNSUPPORTED;          // a chunk is the first or the last of its line, not both
    p.rev = rev;
    p.M = fc.M;
    p.D = fc.D[dir];
    const bool segd = p.seg.nseg > 1;
    const bool rot = env_switch("PBX_YZ_ROT", true) && !segd && !zo.open && !anyT && (dir == 1 || p.G == 1) && p.T >= 2 &&
                     g.nx % XWT == 0;
    CUtensorMap m0, m1;
    if (!make_map_yz(&m0, in0, g, p, rot) || !make_map_yz(&m1, in1, g, p, rot)) return PBX_ERR_UNSUPPORTED;
    const size_t smem = sizeof(YZShared);
    int grid = sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    const bool dot = dir == 2 && pvec && partials;
    const bool fuse = tail && tail->on && dot && !segd && !anyT;
    RedTail t;
    if (fuse) {
        t = *tail;
        t.part = partials;
        t.cnt = t.stride = p.ntx8 * p.ngt;   // one partial sum per 8-wide sub-tile (fast_zpass_max_partials)
        t.narr = 1;
        if (tail_used) *tail_used = true;
    }
#define PBX_YZ(Z, SL, SG, RO, AN, FU, DO) \
    launch_yz<Z, SL, SG, RO, AN, FU, DO>(s, grid, smem, p, (Z) ? zo : ZOpen(), m0, m1, out0, (Z) ? nullptr : out1, (Z) ? pvec : nullptr, \
                                         (Z) ? partials : nullptr, t)
    int rc;
    if (dir == 1) {
        rc = anyT ? PBX_YZ(false, false, false, false, true, false, false)
           : rot  ? PBX_YZ(false, false, false, true, false, false, false)
           : segd ? PBX_YZ(false, false, true, false, false, false, false)
                  : PBX_YZ(false, false, false, false, false, false, false);
    } else if (dot) {
        rc = fuse ? (zo.open ? PBX_YZ(true, true, false, false, false, true, true)
                     : rot   ? PBX_YZ(true, false, false, true, false, true, true)
                             : PBX_YZ(true, false, false, false, false, true, true))
           : anyT ? PBX_YZ(true, false, false, false, true, false, true)
           : rot  ? PBX_YZ(true, false, false, true, false, false, true)
           : zo.open ? PBX_YZ(true, true, false, false, false, false, true)
           : segd ? PBX_YZ(true, false, true, false, false, false, true)
                  : PBX_YZ(true, false, false, false, false, false, true);
    } else {
        rc = anyT ? PBX_YZ(true, false, false, false, true, false, false)
           : rot  ? PBX_YZ(true, false, false, true, false, false, false)
           : zo.open ? PBX_YZ(true, true, false, false, false, false, false)
           : segd ? PBX_YZ(true, false, true, false, false, false, false)
                  : PBX_YZ(true, false, false, false, false, false, false);
    }
#undef PBX_YZ
    PBX_TRY(rc);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// one compact line operator along dir with the TMA-pipelined kernels; PBX_ERR_UNSUPPORTED: use the
// generic kernel (PBX_LINEOP_TMA=0, slab, unsupported shape).  out2 (y, z only):
// the other kind of operator (interpolation <-> derivative) on the same input goes there, in the same
// launch.
int fast_line_op_tma(cudaStream_t s, const Brick &g, int dir, OpKind kind, int stagger, double dx,
                     const double *in, double *out, const double *addend, long long *launches, double *out2)
{
    if (!env_switch("PBX_LINEOP_TMA", true) || !encode_fn()) return PBX_ERR_UNSUPPORTED;
    if (out2 && (dir == 0 || addend)) return PBX_ERR_UNSUPPORTED;
    const lineop::LineOp op = lineop::make_line_op(kind, stagger, dx);
    const lineop::LineOp op2 = lineop::make_line_op(kind == OP_DERIV ? OP_INTERP : OP_DERIV, stagger, dx);
    static std::atomic<bool> attr_set[64];
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<false, false>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<true, false>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<false, true>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<true, true>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<false, false, true>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_kernel<false, true, true>, a, (int)sizeof(LYZShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_x_tma_kernel<false>, a, (int)sizeof(LXShared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_x_tma_kernel<true>, a, (int)sizeof(LXShared)));
        attr_set[dev_ & 63] = true;
    }
    if (dir == 0) {
        const int T = g.nx / LC;
        const size_t nchunks = g.N() / LC;
        if (addend || g.nx % LC || T > NT || (T & (T - 1)) || nchunks > 0x7fffffffull) return PBX_ERR_UNSUPPORTED;
        LXT p;
        p.op = op;
        p.T = T;
        p.ntiles = (int)((nchunks + NT - 1) / NT);
        CUtensorMap mi, mo;
        if (!make_map_x(&mi, in, nchunks) || !make_map_x(&mo, out, nchunks)) return PBX_ERR_UNSUPPORTED;
        int grid = 2 * sm_count();
        if (grid > p.ntiles) grid = p.ntiles;
        if (T > 32)
            lineop_x_tma_kernel<true><<<grid, NT, sizeof(LXShared), s>>>(p, mi, mo);
        else
            lineop_x_tma_kernel<false><<<grid, NT, sizeof(LXShared), s>>>(p, mi, mo);
    } else {
        YZT p;
        if (!yz_geometry_tma(g, dir, &p)) return PBX_ERR_UNSUPPORTED;
        const bool segd = p.seg.nseg > 1;
        p.rev = 0;
        CUtensorMap m;
        if (!make_map_yz(&m, in, g, p)) return PBX_ERR_UNSUPPORTED;
        int grid = sm_count();
        if (grid > p.ntiles) grid = p.ntiles;
        const size_t sm = sizeof(LYZShared);
        if (out2 && segd)
            lineop_yz_tma_kernel<false, true, true><<<grid, NTHR_YZ, sm, s>>>(p, op, m, nullptr, out, op2, out2);
        else if (out2)
            lineop_yz_tma_kernel<false, false, true><<<grid, NTHR_YZ, sm, s>>>(p, op, m, nullptr, out, op2, out2);
        else if (addend && segd)
            lineop_yz_tma_kernel<true, true><<<grid, NTHR_YZ, sm, s>>>(p, op, m, addend, out, op2, nullptr);
        else if (addend)
            lineop_yz_tma_kernel<true, false><<<grid, NTHR_YZ, sm, s>>>(p, op, m, addend, out, op2, nullptr);
        else if (segd)
            lineop_yz_tma_kernel<false, true><<<grid, NTHR_YZ, sm, s>>>(p, op, m, nullptr, out, op2, nullptr);
        else
            lineop_yz_tma_kernel<false, false><<<grid, NTHR_YZ, sm, s>>>(p, op, m, nullptr, out, op2, nullptr);
    }
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// out = kindA(inA) + kindB(inB) along y or z in one launch; PBX_ERR_UNSUPPORTED: the caller uses the generic kernels
int fast_line_op_sum_tma(cudaStream_t s, const Brick &g, int dir, OpKind kindA, OpKind kindB, int stagger,
                         double dx, const double *inA, const double *inB, double *out, long long *launches)
{
    if (!env_switch("PBX_LINEOP_TMA", true) || !encode_fn() || dir == 0) return PBX_ERR_UNSUPPORTED;
    if (out == inA || out == inB ||
        ((reinterpret_cast<uintptr_t>(inA) | reinterpret_cast<uintptr_t>(inB) | reinterpret_cast<uintptr_t>(out)) & 15))
        return PBX_ERR_UNSUPPORTED;
    YZT p;
    if (!yz_geometry_tma(g, dir, &p)) return PBX_ERR_UNSUPPORTED;
    p.rev = 0;
    const lineop::LineOp opA = lineop::make_line_op(kindA, stagger, dx), opB = lineop::make_line_op(kindB, stagger, dx);
    CUtensorMap ma, mb;
    if (!make_map_yz(&ma, inA, g, p) || !make_map_yz(&mb, inB, g, p)) return PBX_ERR_UNSUPPORTED;
    static std::atomic<bool> attr_set[64];
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_sum_kernel<false>, a, (int)sizeof(LYZ2Shared)));
        PBX_CUDA(cudaFuncSetAttribute(lineop_yz_tma_sum_kernel<true>, a, (int)sizeof(LYZ2Shared)));
        attr_set[dev_ & 63] = true;
    }
    int grid = sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    if (p.seg.nseg > 1)
        lineop_yz_tma_sum_kernel<true><<<grid, NTHR_YZ, sizeof(LYZ2Shared), s>>>(p, opA, opB, ma, mb, out);
    else
        lineop_yz_tma_sum_kernel<false><<<grid, NTHR_YZ, sizeof(LYZ2Shared), s>>>(p, opA, opB, ma, mb, out);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
