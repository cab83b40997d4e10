// pbx_dist.cu -- z-slab decomposition over an NCCL communicator.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library loads on machines
// without NCCL and shares the copy a host process (e.g. PyTorch) has already loaded.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "pbx_internal.h"

namespace pbx {

namespace {

// the handful of NCCL entry points used (signatures from nccl.h 2.27; ABI-stable since 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*CommCount)(const ncclComm_t, int *) = nullptr;
    int (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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
    double *buf = nullptr;        // 6 moment arrays [DIST_RMAX][nlines]
    double *send_up, *send_dn, *recv_lo, *recv_up, *self_a, *self_b;
    long long nlines = 0;
    int lower = 0, upper = 0;
};

namespace {

// moments of the boundary planes of the two z-pass inputs: one thread per z line, coalesced in x
__global__ void __launch_bounds__(128)
k_moments(long long nlines, int nzl, int ncs, const double *__restrict__ C,
          const double *__restrict__ D, const double *__restrict__ VAnbM,
          const double *__restrict__ VAnbD, const double *__restrict__ VAsM,
          const double *__restrict__ VAsD, const double *__restrict__ VBnbM,
          const double *__restrict__ VBnbD, const double *__restrict__ VBsM,
          const double *__restrict__ VBsD, double *__restrict__ send_up,
          double *__restrict__ send_dn, double *__restrict__ self_a, double *__restrict__ self_b)
{
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    double up[DIST_RMAX], dn[DIST_RMAX], sa[DIST_RMAX], sb[DIST_RMAX];
#pragma unroll
    for (int a = 0; a < DIST_RMAX; ++a) up[a] = dn[a] = sa[a] = sb[a] = 0.0;
    // bottom planes: neighbour columns of the lower rank's top block (B), own columns of my block A
    for (int k = 0; k < DIST_NB; ++k) {
        const double c = __ldg(C + (long long)k * nlines + l), d = __ldg(D + (long long)k * nlines + l);
#pragma unroll
        for (int a = 0; a < DIST_RMAX; ++a)
            dn[a] = fma(__ldg(VBnbM + k * DIST_RMAX + a), c, fma(__ldg(VBnbD + k * DIST_RMAX + a), d, dn[a]));
        if (k < ncs) {
#pragma unroll
            for (int a = 0; a < DIST_RMAX; ++a)
                sa[a] = fma(__ldg(VAsM + k * DIST_RMAX + a), c, fma(__ldg(VAsD + k * DIST_RMAX + a), d, sa[a]));
        }
    }
    // top planes: neighbour columns of the upper rank's bottom block (A), own columns of my block B
    for (int j = 0; j < DIST_NB; ++j) {
        const int k = nzl - DIST_NB + j;
        const double c = __ldg(C + (long long)k * nlines + l), d = __ldg(D + (long long)k * nlines + l);
#pragma unroll
        for (int a = 0; a < DIST_RMAX; ++a)
            up[a] = fma(__ldg(VAnbM + j * DIST_RMAX + a), c, fma(__ldg(VAnbD + j * DIST_RMAX + a), d, up[a]));
        const int js = k - (nzl - ncs);
        if (js >= 0) {
#pragma unroll
            for (int a = 0; a < DIST_RMAX; ++a)
                sb[a] = fma(__ldg(VBsM + js * DIST_RMAX + a), c, fma(__ldg(VBsD + js * DIST_RMAX + a), d, sb[a]));
        }
    }
#pragma unroll
    for (int a = 0; a < DIST_RMAX; ++a) {
        send_up[a * nlines + l] = up[a];
        send_dn[a * nlines + l] = dn[a];
        self_a[a * nlines + l] = sa[a];
        self_b[a * nlines + l] = sb[a];
    }
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
    PBX_CUDA(cudaMalloc(&d->buf, 6 * per * sizeof(double)));
    PBX_CUDA(cudaMemset(d->buf, 0, 6 * per * sizeof(double)));
    d->send_up = d->buf;
    d->send_dn = d->buf + per;
    d->recv_lo = d->buf + 2 * per;
    d->recv_up = d->buf + 3 * per;
    d->self_a = d->buf + 4 * per;
    d->self_b = d->buf + 5 * per;
    d->lower = (rank + nranks - 1) % nranks;
    d->upper = (rank + 1) % nranks;
    return PBX_OK;
}

int dist_attach(pbx_handle_s *h)
{
    PBX_TRY(nccl_load());
    int n = 1, r = 0;
    PBX_NCCL(g_nccl.CommCount((ncclComm_t)h->comm, &n));
    PBX_NCCL(g_nccl.CommUserRank((ncclComm_t)h->comm, &r));
    return dist_setup(h, r, n);
}

void dist_free(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    if (!d) return;
    if (d->d_tab) cudaFree(d->d_tab);
    if (d->buf) cudaFree(d->buf);
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
    const unsigned nb = (unsigned)((d->nlines + 127) / 128);
    k_moments<<<nb, 128, 0, h->stream>>>(d->nlines, h->nz, d->tab.ncs, S[0], S[1], t + d->oVnbM[0],
                                         t + d->oVnbD[0], t + d->oVsM[0], t + d->oVsD[0],
                                         t + d->oVnbM[1], t + d->oVnbD[1], t + d->oVsM[1],
                                         t + d->oVsD[1], d->send_up, d->send_dn, d->self_a,
                                         d->self_b);
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
    zo.UA = d->d_tab + d->oU[0];
    zo.UB = d->d_tab + d->oU[1];
    zo.mA0 = d->recv_lo;
    zo.mA1 = d->self_a;
    zo.mB0 = d->recv_up;
    zo.mB1 = d->self_b;
    zo.nlines = d->nlines;
    return fast_pass(h, 2, S[0], S[1], out, nullptr, p, partials, &zo);
}

static int dist_exchange_nccl(pbx_handle_s *h)
{
    DistState *d = (DistState *)h->dist;
    const size_t cnt = (size_t)DIST_RMAX * d->nlines;
    ncclComm_t c = (ncclComm_t)h->comm;
    PBX_NCCL(g_nccl.GroupStart());
    PBX_NCCL(g_nccl.Send(d->send_up, cnt, ncclFloat64, d->upper, c, h->stream));
    PBX_NCCL(g_nccl.Send(d->send_dn, cnt, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_lo, cnt, ncclFloat64, d->lower, c, h->stream));
    PBX_NCCL(g_nccl.Recv(d->recv_up, cnt, ncclFloat64, d->upper, c, h->stream));
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
    PBX_TRY(dist_exchange_nccl(h));
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
        PBX_CUDA(cudaMemcpy(up->recv_lo, d->send_up, by, cudaMemcpyDeviceToDevice));
        PBX_CUDA(cudaMemcpy(lo->recv_up, d->send_dn, by, cudaMemcpyDeviceToDevice));
    }
    return PBX_OK;
}

int pbx_comm_destroy(void *comm)
{
    if (!comm) return PBX_OK;
    PBX_TRY(nccl_load());
    PBX_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return PBX_OK;
}

}  // extern "C"
