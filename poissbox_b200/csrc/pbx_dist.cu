// pbx_dist.cu -- z-slab decomposition over an NCCL communicator.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library loads on machines
// without NCCL and shares the copy a host process (e.g. PyTorch) has already loaded.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pbx_internal.h"
#include "pbx_peer.cuh"

namespace pbx {

namespace {

// the handful of NCCL entry points used (signatures from nccl.h 2.27; ABI-stable since 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*CommCount)(const ncclComm_t, int *) = nullptr;
    int (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

Nccl g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return PBX_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        set_last_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        return PBX_ERR_NCCL;
    }
#define SYM(field, name)                                                       \
    do {                                                                       \
        *(void **)(&g_nccl.field) = dlsym(lib, name);                          \
        if (!g_nccl.field) {                                                   \
            set_last_error(std::string("libnccl lacks ") + name);              \
            return PBX_ERR_NCCL;                                               \
        }                                                                      \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(CommCount, "ncclCommCount");
    SYM(CommUserRank, "ncclCommUserRank");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return PBX_OK;
}

int nccl_fail(int e, const char *what)
{
    set_last_error(std::string("NCCL error in ") + what + ": " +
                   (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?"));
    return PBX_ERR_NCCL;
}

#define PBX_NCCL(call)                                               \
    do {                                                             \
        int e__ = (call);                                            \
        if (e__ != ncclSuccess) return nccl_fail(e__, #call);        \
    } while (0)

}  // namespace

// ------------------------------------------------------------------------------------------------
// z-slab state of a handle
// ------------------------------------------------------------------------------------------------
struct DistState {
    double *buf = nullptr;        // local send arrays (used when there is no peer mapping)
    double *send_up, *send_dn;    // [DIST_MSG][nlines]
    // receive buffer: 2 parities x (recv_lo, recv_up), then the peer board; shared by cudaIpc or
    // provided by the host (pbx_slab_link_peers)
    double *rbuf = nullptr;
    bool rbuf_owned = true;
    double *recv_lo[2], *recv_up[2];
    // mappings of the other ranks' receive buffers (cudaIpc over NVLink); nullptr: not mapped
    void *peer_map[PEER_MAXR] = {nullptr};
    double *peer_up_recv_lo[2] = {nullptr, nullptr};   // where my "up" message lands in the upper rank
    double *peer_lo_recv_up[2] = {nullptr, nullptr};   // where my "down" message lands in the lower rank
    // peer boards of ALL ranks mapped: barrier and all-reduce without NCCL (PeerBoard, pbx_internal.h)
    bool peer_sync = false;
    PeerLinks links;
    unsigned long long ar_seq = 0, bar_seq = 0;
    double *sync_word = nullptr;  // device scalar all-reduced as the inter-rank barrier
    unsigned long long epoch = 0; // MatMult counter: parity selects the receive arrays
    long long nlines = 0;
    int lower = 0, upper = 0;
    ZOpen zo;                     // constants of the slab z pass
    // all-gather area of the multigrid preconditioner (mg_slab_plan): 2 parities x gather_cap doubles
    // behind the board
    size_t gather_cap = 0;
    unsigned long long gather_seq = 0;
    size_t per() const { return (size_t)DIST_MSG * (size_t)nlines; }
    size_t board_offset() const { return (4 * per() * sizeof(double) + 127) & ~(size_t)127; }
    size_t gather_offset() const { return (board_offset() + sizeof(PeerBoard) + 127) & ~(size_t)127; }
    size_t recv_bytes() const { return gather_offset() + 2 * gather_cap * sizeof(double); }
    bool peer_stores() const { return peer_up_recv_lo[0] != nullptr; }
};

namespace {

constexpr int BM = 48;   // boundary planes swept for the interpolation recursion (r^48 * 48 < 1e-21)
constexpr int BD = 24;   // ... for the derivative recursion (0.148^24 * 24 < 1e-18)

// zero-halo derivative stencil at window position i of w[0..n): points outside the window are 0
__device__ __forceinline__ double sd_at(const CompositeCoef &D, const double *w, int n, int i)
{
    auto at = [&](int k) { return (k < 0 || k >= n) ? 0.0 : w[k]; };
    const double f0 = w[i];
    const double d1 = fma(-2.0, f0, at(i - 1) + at(i + 1));
    const double d2 = fma(-2.0, f0, at(i - 2) + at(i + 2));
    const double d3 = fma(-2.0, f0, at(i - 3) + at(i + 3));
    return fma(D.c3, d3, fma(D.c2, d2, D.c1 * d1));
}

// Boundary sweep of the two z-pass inputs: what each neighbour needs from my slab (ZOpen in
// pbx_internal.h lists the nine numbers of either message).  One thread per z line, coalesced in
// x; blockIdx.y = 0 sweeps my bottom planes (message to the lower rank), 1 my top planes (to the
// upper rank).  About 3 flops per plane for the interpolation part, 12 for the derivative part.
__global__ void __launch_bounds__(128)
k_boundary(long long nlines, int nzl, const __grid_constant__ CompositeCoef M,
           const __grid_constant__ CompositeCoef D, const double *__restrict__ C,
           const double *__restrict__ Dd, double *__restrict__ msg_dn, double *__restrict__ msg_up)
{
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    if (blockIdx.y == 0) {
        // ---- bottom planes: moments sum r^j u_j, sum j r^j u_j by Horner from plane BM-1 down
        double P = 0.0, Q = 0.0;
        double c0 = 0.0, c1 = 0.0;
        for (int j = BM - 1; j >= 0; --j) {
            const double c = __ldg(C + (long long)j * nlines + l);
            Q = M.r * (Q + P);
            P = fma(M.r, P, c);
            if (j == 1) c1 = c;
            if (j == 0) c0 = c;
        }
        double w[BD + 3];
#pragma unroll
        for (int j = 0; j < BD + 3; ++j) w[j] = __ldg(Dd + (long long)j * nlines + l);
        double PD = 0.0, QD = 0.0;
#pragma unroll
        for (int j = BD - 1; j >= 0; --j) {
            const double sj = sd_at(D, w, BD + 3, j);
            QD = D.r * (QD + PD);
            PD = fma(D.r, PD, sj);
        }
        msg_dn[0 * nlines + l] = P;
        msg_dn[1 * nlines + l] = Q;
        msg_dn[2 * nlines + l] = PD;
        msg_dn[3 * nlines + l] = QD;
        msg_dn[4 * nlines + l] = w[0];
        msg_dn[5 * nlines + l] = w[1];
        msg_dn[6 * nlines + l] = w[2];
        msg_dn[7 * nlines + l] = c0;
        msg_dn[8 * nlines + l] = c1;
    } else {
        // ---- top planes: causal double recursion from zero state, BM (BD) planes below the top
        double y = 0.0, z = 0.0, z1 = 0.0, z2 = 0.0;
        for (int j = nzl - BM; j < nzl; ++j) {
            const double c = __ldg(C + (long long)j * nlines + l);
            y = fma(M.r, y, c);
            z2 = z1;
            z1 = z;
            z = fma(M.r, z, y);
        }
        double w[BD + 3];
#pragma unroll
        for (int j = 0; j < BD + 3; ++j) w[j] = __ldg(Dd + (long long)(nzl - BD - 3 + j) * nlines + l);
        double yD = 0.0, zD = 0.0;
#pragma unroll
        for (int j = 3; j < BD + 3; ++j) {
            const double sj = sd_at(D, w, BD + 3, j);
            yD = fma(D.r, yD, sj);
            zD = fma(D.r, zD, yD);
        }
        msg_up[0 * nlines + l] = y;
        msg_up[1 * nlines + l] = z;
        msg_up[2 * nlines + l] = z1;
        msg_up[3 * nlines + l] = z2;
        msg_up[4 * nlines + l] = yD;
        msg_up[5 * nlines + l] = zD;
        msg_up[6 * nlines + l] = w[BD + 2];
        msg_up[7 * nlines + l] = w[BD + 1];
        msg_up[8 * nlines + l] = w[BD];
    }
}


// Thin slabs (nzl < 2 BM, e.g. 64 planes per rank on eight GPUs): the bottom and the top window of
// the interpolation input overlap, and k_boundary would read the planes in the overlap twice.  Here one
// thread per z line walks the planes of C ONCE, upwards, feeding both the bottom moments (running
// power of r instead of Horner's rule) and the top recursion: 64 planes read instead of 96.  Same
// numbers as k_boundary to rounding; the derivative part is unchanged.
// DOWN: the walk over C goes from the top plane to the bottom one -- the direction that starts on the planes the y
// pass wrote last when it ran front to back (inside the CG), i.e. on what is still in the L2.  The top state is then
// the four moment sums  y = sum r^m c(top-m),  z = sum (m+1) r^m c(top-m),  z' = sum m r^(m-1) c(top-m),
// z'' = sum (m-1) r^(m-2) c(top-m)  over the same BM planes the recursion starts from zero state on, and the
// bottom moments come by Horner's rule as in k_boundary: the same numbers to rounding.
template <bool DOWN>
__global__ void __launch_bounds__(128)
k_boundary_thin(long long nlines, int nzl, const __grid_constant__ CompositeCoef M,
                const __grid_constant__ CompositeCoef D, const double *__restrict__ C,
                const double *__restrict__ Dd, double *__restrict__ msg_dn, double *__restrict__ msg_up)
{
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    // Order of the work: the messages travel over NVLink (2 x 9 numbers per line, 38 MB per apply at 512^2 lines)
    // and the link, not HBM, bounds this kernel when they all leave at the end of the sweep (round 2: 52 us on one
    // GPU, 94 us with the peer stores).  So everything that does not need the long walk over C goes out FIRST --
    // the derivative part (27 planes either side) and the raw planes, 12 of the 18 numbers -- and is in flight
    // while the planes of C stream in; what the walk completes half-way follows as soon as it is complete.
    auto d_bottom = [&]() {
        double w[BD + 3];
#pragma unroll
        for (int j = 0; j < BD + 3; ++j) w[j] = __ldg(Dd + (long long)j * nlines + l);
        double PD = 0.0, QD = 0.0;
#pragma unroll
        for (int j = BD - 1; j >= 0; --j) {
            const double sj = sd_at(D, w, BD + 3, j);
            QD = D.r * (QD + PD);
            PD = fma(D.r, PD, sj);
        }
        msg_dn[2 * nlines + l] = PD;
        msg_dn[3 * nlines + l] = QD;
        msg_dn[4 * nlines + l] = w[0];
        msg_dn[5 * nlines + l] = w[1];
        msg_dn[6 * nlines + l] = w[2];
    };
    auto d_top = [&]() {
        double w[BD + 3];
#pragma unroll
        for (int j = 0; j < BD + 3; ++j) w[j] = __ldg(Dd + (long long)(nzl - BD - 3 + j) * nlines + l);
        double yD = 0.0, zD = 0.0;
#pragma unroll
        for (int j = 3; j < BD + 3; ++j) {
            const double sj = sd_at(D, w, BD + 3, j);
            yD = fma(D.r, yD, sj);
            zD = fma(D.r, zD, yD);
        }
        msg_up[4 * nlines + l] = yD;
        msg_up[5 * nlines + l] = zD;
        msg_up[6 * nlines + l] = w[BD + 2];
        msg_up[7 * nlines + l] = w[BD + 1];
        msg_up[8 * nlines + l] = w[BD];
    };
    if (DOWN) {
        d_top();
        d_bottom();
    } else {
        d_bottom();
        d_top();
    }
    double P = 0.0, Q = 0.0, rj = 1.0;
    double y = 0.0, z = 0.0, z1 = 0.0, z2 = 0.0;
    const int top0 = nzl - BM;
    // planes in blocks of eight, all loads of a block issued before its (serial) recurrences: one load
    // in flight per thread left the sweep latency-bound (65 us for 236 MiB on a 64-plane slab of 512^2 lines)
    constexpr int KB = 8;
    static_assert(BM % KB == 0, "the half-way messages are stored after a whole block of planes");
    double cb[KB], cn[KB];
    if (!DOWN) {
#pragma unroll
        for (int u = 0; u < KB; ++u) cn[u] = __ldg(C + (long long)u * nlines + l);   // nzl >= 64
        for (int j0 = 0; j0 < nzl; j0 += KB) {
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                cb[u] = cn[u];
                const int jn = j0 + KB + u;
                cn[u] = jn < nzl ? __ldg(C + (long long)jn * nlines + l) : 0.0;
            }
            if (j0 == 0) {
                msg_dn[7 * nlines + l] = cb[0];
                msg_dn[8 * nlines + l] = cb[1];
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const int j = j0 + u;
                const double c = cb[u];
                if (j < nzl) {
                    if (j < BM) {
                        const double t = rj * c;
                        P += t;
                        Q = fma((double)j, t, Q);
                        rj *= M.r;
                    }
                    if (j >= top0) {
                        y = fma(M.r, y, c);
                        z2 = z1;
                        z1 = z;
                        z = fma(M.r, z, y);
                    }
                }
            }
            if (j0 == BM - KB) {
                msg_dn[0 * nlines + l] = P;
                msg_dn[1 * nlines + l] = Q;
            }
        }
        msg_up[0 * nlines + l] = y;
        msg_up[1 * nlines + l] = z;
        msg_up[2 * nlines + l] = z1;
        msg_up[3 * nlines + l] = z2;
    } else {
        // m = nzl - 1 - j counts the planes from the top; rm = r^m, rm1 = r^(m-1), rm2 = r^(m-2)
        double rm = 1.0, rm1 = 0.0, rm2 = 0.0;
#pragma unroll
        for (int u = 0; u < KB; ++u) cn[u] = __ldg(C + (long long)(nzl - 1 - u) * nlines + l);
        for (int m0 = 0; m0 < nzl; m0 += KB) {
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                cb[u] = cn[u];
                const int mn = m0 + KB + u;
                cn[u] = mn < nzl ? __ldg(C + (long long)(nzl - 1 - mn) * nlines + l) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const int m = m0 + u, j = nzl - 1 - m;
                const double c = cb[u];
                if (m < nzl) {
                    if (m < BM) {
                        const double t = rm * c;
                        y += t;
                        z = fma((double)(m + 1), t, z);
                        z1 = fma((double)m, rm1 * c, z1);
                        z2 = fma((double)(m - 1), rm2 * c, z2);      // rm2 = 0 while m < 2
                        rm2 = rm1;
                        rm1 = rm;
                        rm *= M.r;
                    }
                    if (j < BM) {
                        Q = M.r * (Q + P);
                        P = fma(M.r, P, c);
                    }
                    if (j == 1) msg_dn[8 * nlines + l] = c;
                    if (j == 0) msg_dn[7 * nlines + l] = c;
                }
            }
            if (m0 == BM - KB) {
                msg_up[0 * nlines + l] = y;
                msg_up[1 * nlines + l] = z;
                msg_up[2 * nlines + l] = z1;
                msg_up[3 * nlines + l] = z2;
            }
        }
        msg_dn[0 * nlines + l] = P;
        msg_dn[1 * nlines + l] = Q;
    }
}

// Neighbour barrier of the slab exchange over the peer boards: the boundary sweep that precedes this
// kernel on the stream has stored my messages into the neighbours' receive arrays; publish that
// (release at system scope: the stores of the earlier kernel happen before it) and wait until both
// neighbours have published theirs.  One thread.
__global__ void k_peer_barrier(const __grid_constant__ PeerLinks L, unsigned long long seq)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ptx::fence_sys();
    ptx::st_release_sys(&L.board[L.upper]->bar_from_lo, seq);
    ptx::st_release_sys(&L.board[L.lower]->bar_from_up, seq);
    const PeerBoard *mine = L.board[L.rank];
    const long long t0 = ptx::spin_start();
    while (ptx::ld_acquire_sys(&mine->bar_from_lo) < seq || ptx::ld_acquire_sys(&mine->bar_from_up) < seq)
        ptx::spin_pause(t0);
}

// All-reduce (sum) of count <= PEER_VALS doubles over the peer boards, in place.  One CTA of
// PEER_MAXR threads at least; thread r talks to rank r.  (The CG folds the same exchange into its
// reduction kernel, pbx_cg.cu; this stand-alone version serves pbx_allreduce_sum.)
__global__ void k_peer_allreduce(const __grid_constant__ PeerLinks L, unsigned long long seq,
                                 double *__restrict__ v, int count)
{
    __shared__ double all[PEER_MAXR][PEER_VALS];
    __shared__ double mine[PEER_VALS];
    if ((int)threadIdx.x < count) mine[threadIdx.x] = v[threadIdx.x];
    __syncthreads();
    double res[PEER_VALS];
    peer_exchange_sum(L, seq, mine, count, all, res);
    if (threadIdx.x == 0)
        for (int a = 0; a < count; ++a) v[a] = res[a];
}

}  // namespace

static void close_peer_maps(DistState *d)
{
    for (int r = 0; r < PEER_MAXR; ++r) {
        if (!d->peer_map[r]) continue;
        cudaIpcCloseMemHandle(d->peer_map[r]);
        for (int q = r + 1; q < PEER_MAXR; ++q)
            if (d->peer_map[q] == d->peer_map[r]) d->peer_map[q] = nullptr;
        d->peer_map[r] = nullptr;
    }
    d->peer_sync = false;
    for (int par = 0; par < 2; ++par) d->peer_up_recv_lo[par] = d->peer_lo_recv_up[par] = nullptr;
}

static void set_recv_pointers(DistState *d)
{
    const size_t per = d->per();
    for (int par = 0; par < 2; ++par) {
        d->recv_lo[par] = d->rbuf + (size_t)(2 * par) * per;
        d->recv_up[par] = d->rbuf + (size_t)(2 * par + 1) * per;
    }
}

// bufs[r]: rank r's receive buffer as addressable from this device (bufs[rank] = my own).  With
// the two neighbours given the boundary sweep stores its messages straight into their memory;
// with ALL ranks given and `boards` set the peer boards take over the barrier and the all-reduce.
static void set_peer_pointers(pbx_handle_s *h, DistState *d, void *const *bufs, bool boards)
{
    const size_t per = d->per();
    for (int par = 0; par < 2; ++par) {
        d->peer_up_recv_lo[par] = (double *)bufs[d->upper] + (size_t)(2 * par) * per;
        d->peer_lo_recv_up[par] = (double *)bufs[d->lower] + (size_t)(2 * par + 1) * per;
    }
    bool all = boards && h->nranks <= PEER_MAXR;
    for (int r = 0; r < h->nranks && all; ++r) all = bufs[r] != nullptr;
    d->peer_sync = false;
    if (!all) return;
    PeerLinks &L = d->links;
    for (int r = 0; r < PEER_MAXR; ++r)
        L.board[r] = r < h->nranks ? (PeerBoard *)((char *)bufs[r] + d->board_offset()) : nullptr;
    L.n = h->nranks;
    L.rank = h->rank;
    L.lower = d->lower;
    L.upper = d->upper;
    d->peer_sync = true;
}

int dist_setup(pbx_handle_s *h, int rank, int nranks)
{
    h->rank = rank;
    h->nranks = nranks;
    if (nranks <= 1) return PBX_OK;
    if (!h->fast_ok) {
        set_last_error("the z-slab decomposition needs the FAST schedule (sizes multiples of 16)");
        return PBX_ERR_UNSUPPORTED;
    }
    if (h->nz < 64 || h->nz > SEG_T * LC) {
        set_last_error("z-slab decomposition needs between 64 and 512 planes per rank");
        return PBX_ERR_UNSUPPORTED;
    }
    DistState *d = new DistState();
    h->dist = d;
    d->nlines = (long long)h->nx * h->ny;
    mg_slab_plan(h->nx, h->ny, h->nz, nranks, &d->gather_cap);
    const size_t per = (size_t)DIST_MSG * d->nlines;
    PBX_CUDA(cudaMalloc(&d->buf, 2 * per * sizeof(double)));
    PBX_CUDA(cudaMemset(d->buf, 0, 2 * per * sizeof(double)));
    d->send_up = d->buf;
    d->send_dn = d->buf + per;
    PBX_CUDA(cudaMalloc(&d->rbuf, d->recv_bytes()));
    PBX_CUDA(cudaMemset(d->rbuf, 0, d->recv_bytes()));
    set_recv_pointers(d);
    PBX_CUDA(cudaMalloc(&d->sync_word, sizeof(double)));
    PBX_CUDA(cudaMemset(d->sync_word, 0, sizeof(double)));
    d->lower = (rank + nranks - 1) % nranks;
    d->upper = (rank + 1) % nranks;
    // closed-form responses of the anti-causal state at the first plane above the slab
    // (sums of geometric series in q = r^2; derivation in DESIGN.md section 6)
    ZOpen &zo = d->zo;
    zo.open = 1;
    zo.nlines = d->nlines;
    const CompositeCoef *cc[2] = {&h->fc.M, &h->fc.D[2]};
    for (int f = 0; f < 2; ++f) {
        const double r = cc[f]->r, q = r * r, i1 = 1.0 / (1.0 - q);
        zo.gw[f] = i1 * i1;
        zo.gx0[f] = (1.0 + q) * i1 * i1 * i1;
        zo.kwz[f] = r * i1;
        zo.kwy[f] = r * i1 * i1;
        zo.kxz[f] = r * i1 * i1;
        zo.kxy[f] = r * (1.0 + q) * i1 * i1 * i1;
    }
    zo.rinv = 1.0 / h->fc.M.r;
    return PBX_OK;
}

int dist_attach(pbx_handle_s *h)
{
    PBX_TRY(nccl_load());
    int n = 1, r = 0;
    PBX_NCCL(g_nccl.CommCount((ncclComm_t)h->comm, &n));
    PBX_NCCL(g_nccl.CommUserRank((ncclComm_t)h->comm, &r));
    PBX_TRY(dist_setup(h, r, n));
    if (n <= 1) return PBX_OK;
    // Exchange cudaIpc handles of the receive arrays and map the two neighbours' arrays, so that
    // the moments kernel stores its results straight into the neighbour's memory over NVLink.
    // Any failure leaves the ncclSend/Recv path in place.
    DistState *d = (DistState *)h->dist;
    if (n > PEER_MAXR) return PBX_OK;   // the same decision on every rank
    // Every rank takes part in the AllGather, whatever happened locally: a rank that cannot (or was told
    // not to, PBX_NO_PEER=1) export its buffer sends a zeroed handle with valid = 0, and the peer / NCCL
    // decision is taken from the GATHERED flags only, so that no rank can leave the others waiting in the
    // collective.
    struct IpcSlot {
        cudaIpcMemHandle_t handle;
        long long valid;
    };
    IpcSlot mine;
    memset(&mine, 0, sizeof mine);
    const char *e = getenv("PBX_NO_PEER");
    if (!(e && e[0] == '1')) {
        if (cudaIpcGetMemHandle(&mine.handle, d->rbuf) == cudaSuccess)
            mine.valid = 1;
        else
            cudaGetLastError();
    }
    char *dall = nullptr;
    const size_t hb = sizeof(IpcSlot);
    PBX_CUDA(cudaMalloc(&dall, hb * (size_t)(n + 1)));
    PBX_CUDA(cudaMemcpy(dall + hb * n, &mine, hb, cudaMemcpyHostToDevice));
    int rc = g_nccl.AllGather(dall + hb * n, dall, hb, ncclInt8, (ncclComm_t)h->comm, h->stream);
    std::vector<IpcSlot> slots(n);
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    if (rc == ncclSuccess && ce == cudaSuccess)
        ce = cudaMemcpy(slots.data(), dall, hb * n, cudaMemcpyDeviceToHost);
    cudaFree(dall);
    if (rc != ncclSuccess || ce != cudaSuccess) {
        // a failed collective is not something one rank can recover from on its own: report it
        cudaGetLastError();
        set_last_error("exchange of the cudaIpc handles failed (ncclAllGather)");
        return rc != ncclSuccess ? PBX_ERR_NCCL : PBX_ERR_CUDA;
    }
    std::vector<cudaIpcMemHandle_t> all(n);
    bool all_valid = true;
    for (int q = 0; q < n; ++q) {
        all[q] = slots[q].handle;
        all_valid = all_valid && slots[q].valid == 1;
    }
    if (!all_valid) return PBX_OK;   // identical on every rank: the ncclSend/Recv path stays in place
    // Peer boards: map EVERY rank's buffer, so that the barrier of the exchange and the CG's
    // all-reduces run over the peer boards (no NCCL call inside an iteration); otherwise the two
    // neighbours only (messages by peer stores, a one-word ncclAllReduce as the barrier)
    // default since the 2- and 8-GPU runs of round 2 (8 GPUs, 512^3: exchange 19.2 -> 9.1 us, CG 0.775 -> 0.722 s);
    // PBX_PEER_SYNC=0 keeps the neighbours-only mapping with the NCCL barrier and all-reduces
    const bool want_all = env_switch("PBX_PEER_SYNC", true) && n <= PEER_MAXR;
    bool ok = true;
    for (int r = 0; r < n && ok; ++r) {
        if (r == h->rank) continue;
        if (!want_all && r != d->upper && r != d->lower) continue;
        ok = cudaIpcOpenMemHandle(&d->peer_map[r], all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
    // every rank must take the same path: agree on success with an all-reduce
    double flag = ok ? 0.0 : 1.0;
    PBX_CUDA(cudaMemcpy(d->sync_word, &flag, sizeof flag, cudaMemcpyHostToDevice));
    PBX_NCCL(g_nccl.AllReduce(d->sync_word, d->sync_word, 1, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream));
    PBX_CUDA(cudaMemcpyAsync(&flag, d->sync_word, sizeof flag, cudaMemcpyDeviceToHost, h->stream));
    PBX_CUDA(cudaStreamSynchronize(h->stream));
    cudaGetLastError();
    if (flag != 0.0) {
        close_peer_maps(d);
        cudaGetLastError();
        return PBX_OK;
    }
    void *bufs[PEER_MAXR] = {nullptr};
    for (int r = 0; r < n; ++r) bufs[r] = r == h->rank ? (void *)d->rbuf : d->peer_map[r];
    set_peer_pointers(h, d, bufs, want_all);
    return PBX_OK;
}

void dist_free(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return;
    close_peer_maps(d);
    if (d->buf) cudaFree(d->buf);
    if (d->rbuf && d->rbuf_owned) cudaFree(d->rbuf);
    if (d->sync_word) cudaFree(d->sync_word);
    delete d;
    h->dist = nullptr;
}

bool dist_connected(const pbx_handle_s *h)
{
    const DistState *d = (const DistState *)h->dist;
    return h->comm != nullptr || (d && d->peer_sync);
}

bool dist_peer_next(pbx_handle_s *h, PeerLinks *L, unsigned long long *seq)
{
    DistState *d = (DistState *)h->dist;
    if (!d || !d->peer_sync) return false;
    *L = d->links;
    *seq = ++d->ar_seq;
    return true;
}

// hand back a sequence number that was drawn for a reduction which did not take place (a reduction tail
// that the z pass could not fuse): executed reductions use CONSECUTIVE sequence numbers, which is what the
// ring-safety argument of pbx_peer.cuh (at most two live slots of PEER_RING) rests on
void dist_peer_unget(pbx_handle_s *h, unsigned long long seq)
{
    DistState *d = (DistState *)h->dist;
    if (d && d->peer_sync && d->ar_seq == seq) --d->ar_seq;
}

int dist_allreduce_sum(pbx_handle_s *h, double *dev, int count)
{
    if (h->nranks <= 1) return PBX_OK;
    if (!dist_connected(h)) {
        set_last_error("this slab handle has no communicator (phase-driven handles cannot reduce)");
        return PBX_ERR_UNSUPPORTED;
    }
#ifdef PBX_DEBUG   // timing experiments only: never in a release build (it skips required communication)
    static const bool skip = getenv("PBX_DEBUG_NO_ALLREDUCE") != nullptr;
    if (skip) return PBX_OK;
#endif
    PeerLinks L;
    unsigned long long seq;
    if (count <= PEER_VALS && dist_peer_next(h, &L, &seq)) {
        k_peer_allreduce<<<1, 32, 0, h->stream>>>(L, seq, dev, count);
        ++h->launches;
        PBX_CUDA(cudaGetLastError());
        return PBX_OK;
    }
    if (!h->comm) {
        set_last_error("peer-board all-reduce carries at most PEER_VALS doubles");
        return PBX_ERR_UNSUPPORTED;
    }
    PBX_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->comm,
                              h->stream));
    return PBX_OK;
}

// x and y sweeps (local) and the boundary sweep that produces the two neighbour messages
int dist_phase1(pbx_handle_s *h, const double *f, int in_cg)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return PBX_ERR_ARG;
    // a segmented y pass (ny > 512) cannot run in place: x pass -> S[2], S[3]; y pass -> S[0], S[1]
    const bool yseg = seg_geometry(h->ny / LC).nseg > 1;
    PBX_TRY(ensure_scratch(h, yseg ? 4 : 2));
    double **S = h->scratch;
    double *A = yseg ? S[2] : S[0], *B = yseg ? S[3] : S[1];
    // tile order and the L2 as in lapl_fast: every pass starts where its producer has just finished (inside the CG the
    // p-update wrote the input front to back).  A 64-plane slab of 512^2 lines is 134 MB per field against 126 MB
    // of L2, so on the slabs of the 8-GPU run this is worth far more than at 512^3 (PBX_SLAB_L2_ORDER=0: all
    // passes front to back, as in round 1)
    const bool ordered = env_switch("PBX_SLAB_L2_ORDER", true);
    const int xrev = (ordered && in_cg) ? 1 : 0, yrev = ordered ? 1 - xrev : 0;
    PBX_TRY(fast_pass(h, 0, f, nullptr, A, B, nullptr, nullptr, nullptr, xrev));
    ++d->epoch;
    const int par = (int)(d->epoch & 1);
    // with peer mappings the messages are stored straight into the neighbours' receive arrays
    double *dst_dn = d->peer_lo_recv_up[par] ? d->peer_lo_recv_up[par] : d->send_dn;
    double *dst_up = d->peer_up_recv_lo[par] ? d->peer_up_recv_lo[par] : d->send_up;
    // (Sending the raw planes of the messages from the y pass instead -- 8 of the 18 numbers per line -- was built
    // and measured on two B200: the y pass's stores over NVLink slowed it by more than the sweep gained, CG iteration
    // 617 against 596 us; profiles/r2_slab_variants.md.)
    PBX_TRY(fast_pass(h, 1, A, B, S[0], S[1], nullptr, nullptr, nullptr, yrev));
    static const bool thin_ok = getenv("PBX_NO_THIN_BOUNDARY") == nullptr;
    if (thin_ok && h->nz < 2 * BM) {
        const unsigned nb = (unsigned)((d->nlines + 127) / 128);
        if (yrev == 0 && ordered)   // the y pass ended on the top planes: walk down from there
            k_boundary_thin<true><<<nb, 128, 0, h->stream>>>(d->nlines, h->nz, h->fc.M, h->fc.D[2], S[0], S[1], dst_dn, dst_up);
        else
            k_boundary_thin<false><<<nb, 128, 0, h->stream>>>(d->nlines, h->nz, h->fc.M, h->fc.D[2], S[0], S[1], dst_dn, dst_up);
    } else {
        dim3 grid((unsigned)((d->nlines + 127) / 128), 2);
        k_boundary<<<grid, 128, 0, h->stream>>>(d->nlines, h->nz, h->fc.M, h->fc.D[2], S[0], S[1], dst_dn,
                                                dst_up);
    }
    ++h->launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// ---- message slots for the line operators of grad / div / interp (pbx_api.cu) --------------------
int dist_begin_epoch(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return PBX_ERR_ARG;
    ++d->epoch;
    return PBX_OK;
}

int dist_line_dst(pbx_handle_s *h, int slot, double **msg_dn, double **msg_up)
{
    DistState *d = (DistState *)h->dist;
    if (!d || slot < 0 || 3 * slot + 3 > DIST_MSG) return PBX_ERR_ARG;
    const int par = (int)(d->epoch & 1);
    const size_t off = (size_t)(3 * slot) * d->nlines;
    *msg_dn = (d->peer_lo_recv_up[par] ? d->peer_lo_recv_up[par] : d->send_dn) + off;
    *msg_up = (d->peer_up_recv_lo[par] ? d->peer_up_recv_lo[par] : d->send_up) + off;
    return PBX_OK;
}

int dist_line_msgs(pbx_handle_s *h, int slot, const double **from_lo, const double **from_up)
{
    DistState *d = (DistState *)h->dist;
    if (!d || slot < 0 || 3 * slot + 3 > DIST_MSG) return PBX_ERR_ARG;
    const int par = (int)(d->epoch & 1);
    const size_t off = (size_t)(3 * slot) * d->nlines;
    *from_lo = d->recv_lo[par] + off;
    *from_up = d->recv_up[par] + off;
    return PBX_OK;
}

// z sweep on the slab, the neighbours' messages of this parity in place
int dist_phase2(pbx_handle_s *h, double *out, const double *p, double *partials)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return PBX_ERR_ARG;
    double **S = h->scratch;
    const int par = (int)(d->epoch & 1);
    ZOpen zo = d->zo;
    zo.from_lo = d->recv_lo[par];
    zo.from_up = d->recv_up[par];
    return fast_pass(h, 2, S[0], S[1], out, nullptr, p, partials, &zo);
}

static int dist_exchange_run(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    if (d->peer_sync) {
        // messages already stored into the neighbours' arrays; flags on the peer boards order the
        // neighbours' boundary sweeps before my z pass
        k_peer_barrier<<<1, 32, 0, h->stream>>>(d->links, ++d->bar_seq);
        ++h->launches;
        PBX_CUDA(cudaGetLastError());
        return PBX_OK;
    }
    ncclComm_t c = (ncclComm_t)h->comm;
    const int par = (int)(d->epoch & 1);
    if (d->peer_stores()) {
        // the boundary sweep has already stored into the neighbours' arrays over NVLink; a
        // one-word all-reduce is the barrier that orders their kernels before my z pass
        PBX_NCCL(g_nccl.AllReduce(d->sync_word, d->sync_word, 1, ncclFloat64, ncclSum, c, h->stream));
        return PBX_OK;
    }
    const size_t cnt = (size_t)DIST_MSG * d->nlines;   // nine planes of nx*ny doubles each way
    PBX_NCCL(g_nccl.GroupStart());
    PBX_NCCL(g_nccl.Send(d->send_up, cnt, ncclFloat64, d->upper, c, h->stream));
    PBX_NCCL(g_nccl.Send(d->send_dn, cnt, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_lo[par], cnt, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_up[par], cnt, ncclFloat64, d->upper, c, h->stream));
    PBX_NCCL(g_nccl.GroupEnd());
    return PBX_OK;
}

int dist_exchange(pbx_handle_s *h)
{
    if (!h->dist || !dist_connected(h)) return PBX_ERR_ARG;
    return dist_exchange_run(h);
}

// One plane each way: my bottom plane of `field` (nz planes of `plane` doubles) becomes the lower
// rank's *hi, my top plane the upper rank's *lo -- the exchange of the star operator (slot 0 of the
// message arrays), for the multigrid levels.
int dist_halo_planes(pbx_handle_s *h, const double *field, size_t plane, int nz, const double **lo,
                     const double **hi)
{
    DistState *d = (DistState *)h->dist;
    if (!d || !dist_connected(h) || plane > (size_t)d->nlines) return PBX_ERR_ARG;
    PBX_TRY(dist_begin_epoch(h));
    double *dn = nullptr, *up = nullptr;
    PBX_TRY(dist_line_dst(h, 0, &dn, &up));
    PBX_CUDA(cudaMemcpyAsync(dn, field, plane * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    PBX_CUDA(cudaMemcpyAsync(up, field + plane * (size_t)(nz - 1), plane * sizeof(double),
                             cudaMemcpyDeviceToDevice, h->stream));
    PBX_TRY(dist_exchange_run(h));
    return dist_line_msgs(h, 0, lo, hi);
}

// The same exchange with the copies left to the producing kernel: begin hands out where my bottom and
// top plane go (the neighbours' buffers of the next round), end is the barrier of that round and returns
// the planes that arrived.
int dist_halo_begin(pbx_handle_s *h, double **dn, double **up)
{
    DistState *d = (DistState *)h->dist;
    if (!d || !dist_connected(h)) return PBX_ERR_ARG;
    PBX_TRY(dist_begin_epoch(h));
    return dist_line_dst(h, 0, dn, up);
}

int dist_halo_end(pbx_handle_s *h, const double **lo, const double **hi)
{
    PBX_TRY(dist_exchange_run(h));
    return dist_line_msgs(h, 0, lo, hi);
}

// All-gather `count` doubles per rank, rank order, into *full (valid until the gather after next).
// Peer boards: every rank copies its part into every rank's gather area (two parities) and an
// all-reduce over the boards is the barrier; otherwise ncclAllGather.
int dist_allgather(pbx_handle_s *h, const double *mine, size_t count, double **full)
{
    DistState *d = (DistState *)h->dist;
    if (!d || !dist_connected(h) || count * (size_t)h->nranks > d->gather_cap) return PBX_ERR_ARG;
    const int par = (int)(++d->gather_seq & 1);
    const size_t off = d->gather_offset() + (size_t)par * d->gather_cap * sizeof(double);
    double *own = (double *)((char *)d->rbuf + off);
    if (d->peer_sync) {
        for (int r = 0; r < h->nranks; ++r) {
            char *base = (char *)d->links.board[r] - d->board_offset();   // rank r's receive buffer
            PBX_CUDA(cudaMemcpyAsync((double *)(base + off) + (size_t)h->rank * count, mine, count * sizeof(double),
                                     cudaMemcpyDeviceToDevice, h->stream));
        }
        PBX_TRY(dist_allreduce_sum(h, d->sync_word, 1));
    } else {
        PBX_NCCL(g_nccl.AllGather(mine, own, count, ncclFloat64, (ncclComm_t)h->comm, h->stream));
    }
    *full = own;
    return PBX_OK;
}

int dist_lapl(pbx_handle_s *h, const double *f, double *out, const double *p, double *partials)
{
    if (!dist_connected(h)) {
        set_last_error("slab handle without a communicator: drive it with pbx_slab_phase1/2");
        return PBX_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) {
        set_last_error("FAST schedule needs 16-byte aligned fields");
        return PBX_ERR_ARG;
    }
    PBX_TRY(dist_phase1(h, f, p != nullptr));
#ifdef PBX_DEBUG   // timing experiments only: never in a release build (it skips required communication)
    static const bool skipx = getenv("PBX_DEBUG_NO_EXCHANGE") != nullptr;
    if (skipx) return dist_phase2(h, out, p, partials);
#endif
    PBX_TRY(dist_exchange_run(h));
    return dist_phase2(h, out, p, partials);
}

}  // namespace pbx

using namespace pbx;

extern "C" {

int pbx_comm_unique_id(void *id128)
{
    if (!id128) return PBX_ERR_ARG;
    PBX_TRY(nccl_load());
    ncclUniqueId id;
    PBX_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return PBX_OK;
}

int pbx_comm_init_rank(const void *id128, int nranks, int rank, int device, void **comm)
{
    if (!id128 || !comm || nranks < 1 || rank < 0 || rank >= nranks) return PBX_ERR_ARG;
    PBX_TRY(nccl_load());
    PBX_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    PBX_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
    *comm = c;
    return PBX_OK;
}

// ---- phase-driven slab handles (single-process emulation of the decomposition, tests) ----------
int pbx_create_slab(int nx, int ny, int nz_local, const double dx[3], int device, int rank,
                    int nranks, pbx_handle *out)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return PBX_ERR_ARG;
    PBX_TRY(pbx_create(nx, ny, nz_local, dx, device, nullptr, out));
    int rc = dist_setup(*out, rank, nranks);
    if (rc != PBX_OK) {
        pbx_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int pbx_slab_phase1(pbx_handle h, const double *f)
{
    if (!h || !f || !h->dist) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return dist_phase1(h, f);
}

int pbx_slab_phase2(pbx_handle h, double *d2f)
{
    if (!h || !d2f || !h->dist) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return dist_phase2(h, d2f, nullptr, nullptr);
}

// the exchange between phase 1 and phase 2 for a ring of slab handles living in ONE process:
// rank r's send_up goes to rank r+1's recv_lo, its send_dn to rank r-1's recv_up
int pbx_slab_exchange_local(pbx_handle *hs, int n)
{
    if (!hs || n < 1) return PBX_ERR_ARG;
    for (int r = 0; r < n; ++r)
        if (!hs[r] || !hs[r]->dist || hs[r]->nranks != n || hs[r]->rank != r) return PBX_ERR_ARG;
    for (int r = 0; r < n; ++r) PBX_CUDA(cudaStreamSynchronize(hs[r]->stream));
    for (int r = 0; r < n; ++r) {
        DistState *d = (DistState *)hs[r]->dist;
        DistState *up = (DistState *)hs[d->upper]->dist, *lo = (DistState *)hs[d->lower]->dist;
        const size_t by = (size_t)DIST_MSG * d->nlines * sizeof(double);
        PBX_CUDA(cudaMemcpy(up->recv_lo[up->epoch & 1], d->send_up, by, cudaMemcpyDeviceToDevice));
        PBX_CUDA(cudaMemcpy(lo->recv_up[lo->epoch & 1], d->send_dn, by, cudaMemcpyDeviceToDevice));
    }
    return PBX_OK;
}

// ---- exchange owned by the host: outgoing messages out, incoming messages in -------------------
int pbx_slab_message_count(pbx_handle h, long long *count)
{
    if (!h || !h->dist || !count) return PBX_ERR_ARG;
    *count = (long long)DIST_MSG * ((DistState *)h->dist)->nlines;
    return PBX_OK;
}

int pbx_slab_get_messages(pbx_handle h, double *up, double *dn)
{
    if (!h || !h->dist || !up || !dn) return PBX_ERR_ARG;
    DistState *d = (DistState *)h->dist;
    if (d->peer_stores()) {
        set_last_error("this handle stores its messages straight into the neighbours' memory");
        return PBX_ERR_UNSUPPORTED;
    }
    PBX_CUDA(cudaSetDevice(h->device));
    const size_t by = (size_t)DIST_MSG * d->nlines * sizeof(double);
    PBX_CUDA(cudaMemcpyAsync(up, d->send_up, by, cudaMemcpyDeviceToDevice, h->stream));
    PBX_CUDA(cudaMemcpyAsync(dn, d->send_dn, by, cudaMemcpyDeviceToDevice, h->stream));
    PBX_CUDA(cudaStreamSynchronize(h->stream));
    return PBX_OK;
}

int pbx_slab_put_messages(pbx_handle h, const double *from_lo, const double *from_up)
{
    if (!h || !h->dist || !from_lo || !from_up) return PBX_ERR_ARG;
    DistState *d = (DistState *)h->dist;
    PBX_CUDA(cudaSetDevice(h->device));
    const size_t by = (size_t)DIST_MSG * d->nlines * sizeof(double);
    const int par = (int)(d->epoch & 1);
    PBX_CUDA(cudaMemcpyAsync(d->recv_lo[par], from_lo, by, cudaMemcpyDeviceToDevice, h->stream));
    PBX_CUDA(cudaMemcpyAsync(d->recv_up[par], from_up, by, cudaMemcpyDeviceToDevice, h->stream));
    return PBX_OK;
}

// ---- peer boards linked by the host --------------------------------------------------------------
int pbx_slab_recv_bytes(pbx_handle h, size_t *bytes)
{
    if (!h || !h->dist || !bytes) return PBX_ERR_ARG;
    *bytes = ((DistState *)h->dist)->recv_bytes();
    return PBX_OK;
}

int pbx_slab_recv_buffer(pbx_handle h, void **buf)
{
    if (!h || !h->dist || !buf) return PBX_ERR_ARG;
    *buf = ((DistState *)h->dist)->rbuf;
    return PBX_OK;
}

int pbx_peer_sync_active(pbx_handle h)
{
    if (!h || !h->dist) return -PBX_ERR_ARG;
    return ((DistState *)h->dist)->peer_sync ? 1 : 0;
}

int pbx_slab_link_peers(pbx_handle h, void *const *bufs, int n)
{
    if (!h || !h->dist || !bufs || n != h->nranks || n < 2) return PBX_ERR_ARG;
    if (n > PEER_MAXR) {
        set_last_error("peer boards serve at most 16 ranks");
        return PBX_ERR_UNSUPPORTED;
    }
    for (int r = 0; r < n; ++r)
        if (!bufs[r] || (reinterpret_cast<uintptr_t>(bufs[r]) & 127)) return PBX_ERR_ARG;
    DistState *d = (DistState *)h->dist;
    PBX_CUDA(cudaSetDevice(h->device));
    if (d->peer_stores() && d->peer_map[d->upper]) {
        set_last_error("this handle is already linked to its peers over cudaIpc");
        return PBX_ERR_UNSUPPORTED;
    }
    if (bufs[h->rank] != (void *)d->rbuf) {
        // the host provides my receive buffer (e.g. memory it shares with the other ranks)
        PBX_CUDA(cudaStreamSynchronize(h->stream));
        if (d->rbuf_owned) cudaFree(d->rbuf);
        d->rbuf = (double *)bufs[h->rank];
        d->rbuf_owned = false;
        set_recv_pointers(d);
    }
    set_peer_pointers(h, d, bufs, true);
    return PBX_OK;
}

// the exchange step alone, over the handle's communicator (profiling aid)
int pbx_slab_exchange(pbx_handle h)
{
    if (!h || !h->dist || !dist_connected(h)) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return dist_exchange_run(h);
}

// sum `count` device doubles over the handle's communicator, in place (profiling aid)
int pbx_allreduce_sum(pbx_handle h, double *dev, int count)
{
    if (!h || !dev || count < 1) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return dist_allreduce_sum(h, dev, count);
}

int pbx_comm_destroy(void *comm)
{
    if (!comm) return PBX_OK;
    PBX_TRY(nccl_load());
    PBX_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return PBX_OK;
}

}  // extern "C"
