// pbx_dist.cu -- z-slab decomposition over an NCCL communicator.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library loads on machines
// without NCCL and shares the copy a host process (e.g. PyTorch) has already loaded.
#include <dlfcn.h>

#include <cstring>

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

int dist_attach(pbx_handle_s *h)
{
    PBX_TRY(nccl_load());
    int n = 1, r = 0;
    PBX_NCCL(g_nccl.CommCount((ncclComm_t)h->comm, &n));
    PBX_NCCL(g_nccl.CommUserRank((ncclComm_t)h->comm, &r));
    h->nranks = n;
    h->rank = r;
    if (n > 1 && !h->fast_ok) {
        set_last_error("the z-slab decomposition needs the FAST schedule (sizes multiples of 16)");
        return PBX_ERR_UNSUPPORTED;
    }
    return PBX_OK;
}

void dist_free(pbx_handle_s *h) { (void)h; }

int dist_allreduce_sum(pbx_handle_s *h, double *dev, int count)
{
    if (h->nranks <= 1) return PBX_OK;
    PBX_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->comm,
                              h->stream));
    return PBX_OK;
}

int dist_lapl(pbx_handle_s *h, const double *f, double *out, const double *p, double *partials)
{
    (void)h; (void)f; (void)out; (void)p; (void)partials;
    set_last_error("z-slab Laplacian: not built yet");
    return PBX_ERR_UNSUPPORTED;
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

int pbx_comm_destroy(void *comm)
{
    if (!comm) return PBX_OK;
    PBX_TRY(nccl_load());
    PBX_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return PBX_OK;
}

}  // extern "C"
