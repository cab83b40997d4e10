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
    DistTables tab;
    double *d_tab = nullptr;      // device copy of the tables, see offsets below
    size_t oU[2], oVnbM[2], oVsM[2], oVnbD[2], oVsD[2];
    double *buf = nullptr;        // 4 local moment arrays [DIST_RMAX][nlines]
    double *send_up, *send_dn, *self_a, *self_b;
    double *rbuf = nullptr;       // receive arrays, 2 parities x (recv_lo, recv_up); shared by cudaIpc
    double *recv_lo[2], *recv_up[2];
    // NVLink peer mappings of the neighbours' receive arrays (nullptr: use ncclSend/Recv)
    void *peer_map_up = nullptr, *peer_map_lo = nullptr;
    double *peer_up_recv_lo[2] = {nullptr, nullptr};   // where my send_up lands in the upper rank
    double *peer_lo_recv_up[2] = {nullptr, nullptr};   // where my send_dn lands in the lower rank
    double *sync_word = nullptr;  // device scalar all-reduced as the inter-rank barrier
    unsigned long long epoch = 0; // MatMult counter: parity selects the receive arrays
    long long nlines = 0;
    int lower = 0, upper = 0;
    int rows_nb[2][2], rows_s[2][2];   // [side][M, D]: table rows that matter, from the boundary
    int rows_u[2];                      // rows of U that matter
};

namespace {

// Moments of the boundary planes of the two z-pass inputs.  blockIdx.y selects the boundary
// (0: my bottom planes, 1: my top planes); a thread owns two z lines (x, x + 1) so that every
// table entry fetched from shared memory feeds four FMAs.  For the boundary b the planes are both
// "neighbour columns" of the adjacent rank's block (-> send array) and "own columns" of my block
// (-> self array).  Table rows whose entries are all below 1e-19 of the largest are skipped
// (nM / nD planes for the interpolation / derivative part).
struct MomArgs {
    const double *VnbM, *VnbD, *VsM, *VsD;   // device tables of the side, [plane][DIST_RMAX]
    double *send, *self;                      // [DIST_RMAX][nlines]
    int Rnb, Rs;                              // moments actually used
    int nbM, nbD, nsM, nsD;                   // planes that matter, counted from the boundary
};

__global__ void __launch_bounds__(128)
k_moments(long long nlines, int nzl, int ncs, const double *__restrict__ C,
          const double *__restrict__ D, const __grid_constant__ MomArgs bot,
          const __grid_constant__ MomArgs top)
{
    __shared__ __align__(16) double tab[4][DIST_NB][DIST_RMAX];
    const bool is_top = blockIdx.y == 1;
    const MomArgs &A = is_top ? top : bot;
    for (int i = threadIdx.x; i < DIST_NB * DIST_RMAX; i += blockDim.x) {
        const int k = i / DIST_RMAX;
        (&tab[0][0][0])[i] = A.VnbM[i];
        (&tab[1][0][0])[i] = A.VnbD[i];
        (&tab[2][0][0])[i] = k < ncs ? A.VsM[i] : 0.0;
        (&tab[3][0][0])[i] = k < ncs ? A.VsD[i] : 0.0;
    }
    __syncthreads();
    const long long l = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (l >= nlines) return;
    const bool two = l + 1 < nlines;
    double nb0[DIST_RMAX], nb1[DIST_RMAX], sf0[DIST_RMAX], sf1[DIST_RMAX];
#pragma unroll
    for (int a = 0; a < DIST_RMAX; ++a) nb0[a] = nb1[a] = sf0[a] = sf1[a] = 0.0;
    // j counts planes in the table's own order: for the bottom boundary table row j is plane j;
    // for the top boundary the neighbour table row j is plane nzl-NB+j and the own-column table
    // row j is plane nzl-ncs+j
    const int nmax = DIST_NB;
    for (int j = 0; j < nmax; ++j) {
        const int k = is_top ? nzl - DIST_NB + j : j;
        const double2 c = two ? *reinterpret_cast<const double2 *>(C + (long long)k * nlines + l)
                              : make_double2(C[(long long)k * nlines + l], 0.0);
        const double2 d = two ? *reinterpret_cast<const double2 *>(D + (long long)k * nlines + l)
                              : make_double2(D[(long long)k * nlines + l], 0.0);
        // distance from the boundary decides whether the row still matters
        const int dist = is_top ? DIST_NB - 1 - j : j;
        if (dist < A.nbM) {
#pragma unroll
            for (int a = 0; a < DIST_RMAX; ++a) {
                const double v = tab[0][j][a];
                nb0[a] = fma(v, c.x, nb0[a]);
                nb1[a] = fma(v, c.y, nb1[a]);
            }
        }
        if (dist < A.nbD) {
#pragma unroll
            for (int a = 0; a < DIST_RMAX; ++a) {
                const double v = tab[1][j][a];
                nb0[a] = fma(v, d.x, nb0[a]);
                nb1[a] = fma(v, d.y, nb1[a]);
            }
        }
        const int js = is_top ? k - (nzl - ncs) : j;     // row of the own-column table
        if (js >= 0 && js < ncs) {
            const int ds = is_top ? ncs - 1 - js : js;
            if (ds < A.nsM) {
#pragma unroll
                for (int a = 0; a < DIST_RMAX; ++a) {
                    const double v = tab[2][js][a];
                    sf0[a] = fma(v, c.x, sf0[a]);
                    sf1[a] = fma(v, c.y, sf1[a]);
                }
            }
            if (ds < A.nsD) {
#pragma unroll
                for (int a = 0; a < DIST_RMAX; ++a) {
                    const double v = tab[3][js][a];
                    sf0[a] = fma(v, d.x, sf0[a]);
                    sf1[a] = fma(v, d.y, sf1[a]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < DIST_RMAX; ++a) {
        if (two) {
            *reinterpret_cast<double2 *>(A.send + a * nlines + l) = make_double2(nb0[a], nb1[a]);
            *reinterpret_cast<double2 *>(A.self + a * nlines + l) = make_double2(sf0[a], sf1[a]);
        } else {
            A.send[a * nlines + l] = nb0[a];
            A.self[a * nlines + l] = sf0[a];
        }
    }
}

// number of leading (distance-from-boundary ordered) rows of a table that matter
int rows_that_matter(const std::vector<double> &V, int nrows, bool top_order)
{
    double mx = 0.0;
    for (double v : V) mx = std::max(mx, std::fabs(v));
    int need = 0;
    for (int j = 0; j < nrows; ++j) {
        double rm = 0.0;
        for (int a = 0; a < DIST_RMAX; ++a) rm = std::max(rm, std::fabs(V[(size_t)j * DIST_RMAX + a]));
        const int dist = top_order ? nrows - 1 - j : j;
        if (rm > 1e-19 * mx) need = std::max(need, dist + 1);
    }
    return need;
}

}  // namespace

int dist_setup(pbx_handle_s *h, int rank, int nranks)
{
    h->rank = rank;
    h->nranks = nranks;
    if (nranks <= 1) return PBX_OK;
    if (!h->fast_ok) {
        set_last_error("the z-slab decomposition needs the FAST schedule (sizes multiples of 16)");
        return PBX_ERR_UNSUPPORTED;
    }
    DistState *d = new DistState();
    h->dist = d;
    PBX_TRY(build_dist_tables(h->nz, h->fc.M, h->fc.D[2], &d->tab));
    // pack the tables for the device
    std::vector<double> pk;
    for (int s = 0; s < 2; ++s) {
        const DistSide &S = d->tab.side[s];
        auto add = [&](const std::vector<double> &v, size_t *off) {
            *off = pk.size();
            pk.insert(pk.end(), v.begin(), v.end());
        };
        add(S.U, &d->oU[s]);
        add(S.VnbM, &d->oVnbM[s]);
        add(S.VsM, &d->oVsM[s]);
        add(S.VnbD, &d->oVnbD[s]);
        add(S.VsD, &d->oVsD[s]);
    }
    PBX_CUDA(cudaMalloc(&d->d_tab, pk.size() * sizeof(double)));
    PBX_CUDA(cudaMemcpy(d->d_tab, pk.data(), pk.size() * sizeof(double), cudaMemcpyHostToDevice));
    d->nlines = (long long)h->nx * h->ny;
    const size_t per = (size_t)DIST_RMAX * d->nlines;
    PBX_CUDA(cudaMalloc(&d->buf, 4 * per * sizeof(double)));
    PBX_CUDA(cudaMemset(d->buf, 0, 4 * per * sizeof(double)));
    d->send_up = d->buf;
    d->send_dn = d->buf + per;
    d->self_a = d->buf + 2 * per;
    d->self_b = d->buf + 3 * per;
    PBX_CUDA(cudaMalloc(&d->rbuf, 4 * per * sizeof(double)));
    PBX_CUDA(cudaMemset(d->rbuf, 0, 4 * per * sizeof(double)));
    for (int par = 0; par < 2; ++par) {
        d->recv_lo[par] = d->rbuf + (size_t)(2 * par) * per;
        d->recv_up[par] = d->rbuf + (size_t)(2 * par + 1) * per;
    }
    PBX_CUDA(cudaMalloc(&d->sync_word, sizeof(double)));
    PBX_CUDA(cudaMemset(d->sync_word, 0, sizeof(double)));
    d->lower = (rank + nranks - 1) % nranks;
    d->upper = (rank + 1) % nranks;
    // neighbour columns of the bottom block (side 0) are the lower rank's TOP planes: the row
    // nearest to the boundary is the last one; for the top block (side 1) it is the first one.
    // Own columns: bottom block -> first row nearest, top block -> last row nearest.
    for (int sd = 0; sd < 2; ++sd) {
        const DistSide &S = d->tab.side[sd];
        d->rows_nb[sd][0] = rows_that_matter(S.VnbM, DIST_NB, sd == 0);
        d->rows_nb[sd][1] = rows_that_matter(S.VnbD, DIST_NB, sd == 0);
        d->rows_s[sd][0] = rows_that_matter(S.VsM, d->tab.ncs, sd == 1);
        d->rows_s[sd][1] = rows_that_matter(S.VsD, d->tab.ncs, sd == 1);
        d->rows_u[sd] = rows_that_matter(S.U, d->tab.nrow, sd == 1);
    }
    if ((h->nx & 1) != 0) {
        set_last_error("z-slab decomposition needs an even nx");
        return PBX_ERR_UNSUPPORTED;
    }
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
    const char *e = getenv("PBX_NO_PEER");
    if (e && e[0] == '1') return PBX_OK;
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, d->rbuf) != cudaSuccess) {
        cudaGetLastError();
        return PBX_OK;
    }
    char *dall = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    PBX_CUDA(cudaMalloc(&dall, hb * (size_t)(n + 1)));
    PBX_CUDA(cudaMemcpy(dall + hb * n, &mine, hb, cudaMemcpyHostToDevice));
    int rc = g_nccl.AllGather(dall + hb * n, dall, hb, ncclInt8, (ncclComm_t)h->comm, h->stream);
    std::vector<cudaIpcMemHandle_t> all(n);
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    if (rc == ncclSuccess && ce == cudaSuccess)
        ce = cudaMemcpy(all.data(), dall, hb * n, cudaMemcpyDeviceToHost);
    cudaFree(dall);
    if (rc != ncclSuccess || ce != cudaSuccess) {
        cudaGetLastError();
        return PBX_OK;
    }
    bool ok = cudaIpcOpenMemHandle(&d->peer_map_up, all[d->upper], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (ok) {
        if (d->lower == d->upper)
            d->peer_map_lo = d->peer_map_up;
        else
            ok = cudaIpcOpenMemHandle(&d->peer_map_lo, all[d->lower], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
    // every rank must take the same path: agree on success with an all-reduce
    double flag = ok ? 0.0 : 1.0;
    PBX_CUDA(cudaMemcpy(d->sync_word, &flag, sizeof flag, cudaMemcpyHostToDevice));
    PBX_NCCL(g_nccl.AllReduce(d->sync_word, d->sync_word, 1, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream));
    PBX_CUDA(cudaMemcpyAsync(&flag, d->sync_word, sizeof flag, cudaMemcpyDeviceToHost, h->stream));
    PBX_CUDA(cudaStreamSynchronize(h->stream));
    cudaGetLastError();
    if (flag != 0.0) {
        if (d->peer_map_up) cudaIpcCloseMemHandle(d->peer_map_up);
        if (d->peer_map_lo && d->peer_map_lo != d->peer_map_up) cudaIpcCloseMemHandle(d->peer_map_lo);
        d->peer_map_up = d->peer_map_lo = nullptr;
        cudaGetLastError();
        return PBX_OK;
    }
    const size_t per = (size_t)DIST_RMAX * d->nlines;
    for (int par = 0; par < 2; ++par) {
        d->peer_up_recv_lo[par] = (double *)d->peer_map_up + (size_t)(2 * par) * per;
        d->peer_lo_recv_up[par] = (double *)d->peer_map_lo + (size_t)(2 * par + 1) * per;
    }
    return PBX_OK;
}

void dist_free(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return;
    if (d->peer_map_up) cudaIpcCloseMemHandle(d->peer_map_up);
    if (d->peer_map_lo && d->peer_map_lo != d->peer_map_up) cudaIpcCloseMemHandle(d->peer_map_lo);
    if (d->d_tab) cudaFree(d->d_tab);
    if (d->buf) cudaFree(d->buf);
    if (d->rbuf) cudaFree(d->rbuf);
    if (d->sync_word) cudaFree(d->sync_word);
    delete d;
    h->dist = nullptr;
}

int dist_allreduce_sum(pbx_handle_s *h, double *dev, int count)
{
    if (h->nranks <= 1) return PBX_OK;
    if (!h->comm) {
        set_last_error("this slab handle has no communicator (phase-driven handles cannot reduce)");
        return PBX_ERR_UNSUPPORTED;
    }
    static const bool skip = getenv("PBX_DEBUG_NO_ALLREDUCE") != nullptr;   // timing experiments only
    if (skip) return PBX_OK;
    PBX_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->comm,
                              h->stream));
    return PBX_OK;
}

// x and y sweeps (local) and the moments of the boundary planes of the z-pass inputs
int dist_phase1(pbx_handle_s *h, const double *f)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return PBX_ERR_ARG;
    PBX_TRY(ensure_scratch(h, 2));
    double **S = h->scratch;
    PBX_TRY(fast_pass(h, 0, f, nullptr, S[0], S[1], nullptr, nullptr));
    PBX_TRY(fast_pass(h, 1, S[0], S[1], S[0], S[1], nullptr, nullptr));
    const double *t = d->d_tab;
    // my BOTTOM planes are neighbour columns of the lower rank's top block (side 1) and own
    // columns of my bottom block (side 0); my TOP planes the other way round
    ++d->epoch;
    const int par = (int)(d->epoch & 1);
    // with peer mappings the "send" arrays ARE the neighbours' receive arrays of this parity
    double *dst_dn = d->peer_lo_recv_up[par] ? d->peer_lo_recv_up[par] : d->send_dn;
    double *dst_up = d->peer_up_recv_lo[par] ? d->peer_up_recv_lo[par] : d->send_up;
    MomArgs bot{t + d->oVnbM[1], t + d->oVnbD[1], t + d->oVsM[0], t + d->oVsD[0], dst_dn, d->self_a,
                d->tab.side[1].R, d->tab.side[0].R, d->rows_nb[1][0], d->rows_nb[1][1],
                d->rows_s[0][0], d->rows_s[0][1]};
    MomArgs top{t + d->oVnbM[0], t + d->oVnbD[0], t + d->oVsM[1], t + d->oVsD[1], dst_up, d->self_b,
                d->tab.side[0].R, d->tab.side[1].R, d->rows_nb[0][0], d->rows_nb[0][1],
                d->rows_s[1][0], d->rows_s[1][1]};
    const long long pairs = (d->nlines + 1) / 2;
    dim3 grid((unsigned)((pairs + 127) / 128), 2);
    k_moments<<<grid, 128, 0, h->stream>>>(d->nlines, h->nz, d->tab.ncs, S[0], S[1], bot, top);
    ++h->launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// z sweep on the open slab with the boundary corrections (needs recv_lo / recv_up filled)
int dist_phase2(pbx_handle_s *h, double *out, const double *p, double *partials)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return PBX_ERR_ARG;
    double **S = h->scratch;
    ZOpen zo;
    zo.open = 1;
    zo.nrow = d->tab.nrow;
    zo.nrowA = d->rows_u[0];
    zo.nrowB = d->rows_u[1];
    zo.RA = d->tab.side[0].R;
    zo.RB = d->tab.side[1].R;
    zo.UA = d->d_tab + d->oU[0];
    zo.UB = d->d_tab + d->oU[1];
    const int par = (int)(d->epoch & 1);
    zo.mA0 = d->recv_lo[par];
    zo.mA1 = d->self_a;
    zo.mB0 = d->recv_up[par];
    zo.mB1 = d->self_b;
    zo.nlines = d->nlines;
    return fast_pass(h, 2, S[0], S[1], out, nullptr, p, partials, &zo);
}

static int dist_exchange_nccl(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    ncclComm_t c = (ncclComm_t)h->comm;
    const int par = (int)(d->epoch & 1);
    if (d->peer_map_up) {
        // the moments kernel has already stored into the neighbours' arrays over NVLink; a
        // one-word all-reduce is the barrier that orders their kernels before my z pass
        PBX_NCCL(g_nccl.AllReduce(d->sync_word, d->sync_word, 1, ncclFloat64, ncclSum, c, h->stream));
        return PBX_OK;
    }
    // only the R moments that exist travel: 7 planes of nx*ny doubles up, 5 down
    const size_t cup = (size_t)d->tab.side[0].R * d->nlines, cdn = (size_t)d->tab.side[1].R * d->nlines;
    PBX_NCCL(g_nccl.GroupStart());
    PBX_NCCL(g_nccl.Send(d->send_up, cup, ncclFloat64, d->upper, c, h->stream));
    PBX_NCCL(g_nccl.Send(d->send_dn, cdn, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_lo[par], cup, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_up[par], cdn, ncclFloat64, d->upper, c, h->stream));
    PBX_NCCL(g_nccl.GroupEnd());
    return PBX_OK;
}

int dist_lapl(pbx_handle_s *h, const double *f, double *out, const double *p, double *partials)
{
    if (!h->comm) {
        set_last_error("slab handle without a communicator: drive it with pbx_slab_phase1/2");
        return PBX_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) {
        set_last_error("FAST schedule needs 16-byte aligned fields");
        return PBX_ERR_ARG;
    }
    PBX_TRY(dist_phase1(h, f));
    static const bool skipx = getenv("PBX_DEBUG_NO_EXCHANGE") != nullptr;   // timing experiments only
    if (!skipx) PBX_TRY(dist_exchange_nccl(h));
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
        const size_t by = (size_t)DIST_RMAX * d->nlines * sizeof(double);
        PBX_CUDA(cudaMemcpy(up->recv_lo[up->epoch & 1], d->send_up, by, cudaMemcpyDeviceToDevice));
        PBX_CUDA(cudaMemcpy(lo->recv_up[lo->epoch & 1], d->send_dn, by, cudaMemcpyDeviceToDevice));
    }
    return PBX_OK;
}

// the exchange step alone, over the handle's communicator (profiling aid)
int pbx_slab_exchange(pbx_handle h)
{
    if (!h || !h->dist || !h->comm) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return dist_exchange_nccl(h);
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
