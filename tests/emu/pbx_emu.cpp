// pbx_emu.cpp -- fiber scheduler and fake device of the CPU kernel-logic harness (see
// include/pbx_emu.h).  TEST INFRASTRUCTURE ONLY.
#include <ucontext.h>

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

namespace pbx_emu {

long long tensor_maps_total();
long long tensor_maps3_swizzled_total();

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {

constexpr size_t STACK = 256 * 1024;
constexpr size_t SMEM = 256 * 1024;
constexpr int NBAR = 16;

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    uint3 tid;
    int lin = 0;
    bool done = true;
};

struct Bar {
    int count = 0;
    unsigned gen = 0;
};

std::vector<Fiber *> g_pool;
ucontext_t g_main;
Fiber *g_cur = nullptr;
const std::function<void()> *g_body = nullptr;
unsigned char *g_smem = nullptr;
Bar g_bars[NBAR];
std::vector<Bar> g_warp_bars;
std::vector<uint64_t> g_warp_slots;
int g_nthreads = 0;
unsigned long long g_progress = 0;
long long g_launches = 0;

void fiber_entry()
{
    (*g_body)();
    g_cur->done = true;
    ++g_progress;
    swapcontext(&g_cur->ctx, &g_main);
}

void wait_bar(Bar &b, int count)
{
    const unsigned gen = b.gen;
    ++g_progress;   // an arrival changes the state
    if (++b.count >= count) {
        b.count = 0;
        ++b.gen;
        ++g_progress;
        return;
    }
    while (b.gen == gen) yield();
}

}  // namespace

void die(const char *what)
{
    fprintf(stderr, "pbx_emu: %s (block %u,%u,%u thread %u,%u,%u)\n", what, g_blockIdx.x, g_blockIdx.y,
            g_blockIdx.z, g_threadIdx.x, g_threadIdx.y, g_threadIdx.z);
    abort();
}

unsigned char *dyn_smem() { return g_smem; }
void note_progress() { ++g_progress; }
int linear_tid() { return g_cur ? g_cur->lin : 0; }
long long launches_total() { return g_launches; }

void yield()
{
    if (!g_cur) die("yield outside a kernel");
    swapcontext(&g_cur->ctx, &g_main);
}

void barrier(int id, int count)
{
    if (id < 0 || id >= NBAR) die("named barrier id out of range");
    if (count > g_nthreads) die("barrier expects more threads than the CTA has");
    wait_bar(g_bars[id], count);
}

uint64_t shfl_bits(uint64_t v, int src)
{
    const int lin = g_cur->lin, w = lin >> 5, lane = lin & 31;
    const int live = g_nthreads - 32 * w < 32 ? g_nthreads - 32 * w : 32;
    if (src < 0 || src >= live) src = lane;
    g_warp_slots[(size_t)w * 32 + lane] = v;
    wait_bar(g_warp_bars[2 * w], live);
    const uint64_t r = g_warp_slots[(size_t)w * 32 + src];
    wait_bar(g_warp_bars[2 * w + 1], live);
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
    if (g_cur) die("nested kernel launch");
    const int nthr = (int)(block.x * block.y * block.z);
    if (nthr < 1 || nthr > 1024) die("bad CTA size");
    if (smem_bytes > SMEM) die("dynamic shared memory request too large");
    if (!g_smem) g_smem = (unsigned char *)aligned_alloc(1024, SMEM);
    while ((int)g_pool.size() < nthr) {
        Fiber *f = new Fiber();
        f->stack = (char *)malloc(STACK);
        g_pool.push_back(f);
    }
    ++g_launches;
    g_body = &body;
    g_blockDim = block;
    g_gridDim = grid;
    g_nthreads = nthr;
    const int nwarps = (nthr + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                memset(g_smem, 0xff, SMEM);   // NaN: shared memory is uninitialised at CTA start
                for (auto &b : g_bars) b = Bar();
                g_warp_bars.assign(2 * nwarps, Bar());
                g_warp_slots.assign((size_t)nwarps * 32, 0);
                int lin = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++lin) {
                            Fiber *f = g_pool[lin];
                            f->tid = uint3{tx, ty, tz};
                            f->lin = lin;
                            f->done = false;
                            getcontext(&f->ctx);
                            f->ctx.uc_stack.ss_sp = f->stack;
                            f->ctx.uc_stack.ss_size = STACK;
                            f->ctx.uc_link = nullptr;
                            makecontext(&f->ctx, fiber_entry, 0);
                        }
                // Thread order within a scheduling round.  A fiber runs undisturbed from one switch point
                // (barrier, shuffle, mbarrier wait) to the next, so the order decides which of two
                // unsynchronised accesses between the same barriers happens first: a kernel without
                // such races gives the same bits under every order (PBX_EMU_SCHED = 0 ascending,
                // 1 descending, n >= 2 pseudo-random with seed n, reshuffled every round).
                static const int sched = getenv("PBX_EMU_SCHED") ? atoi(getenv("PBX_EMU_SCHED")) : 0;
                std::vector<int> order(nthr);
                for (int i = 0; i < nthr; ++i) order[i] = sched == 1 ? nthr - 1 - i : i;
                unsigned long long rng = 0x9E3779B97F4A7C15ull * (unsigned long long)(sched + 1) + bx + 131 * by + 7919 * bz;
                int remaining = nthr;
                while (remaining > 0) {
                    const unsigned long long before = g_progress;
                    remaining = 0;
                    if (sched >= 2) {
                        for (int i = nthr - 1; i > 0; --i) {
                            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                            std::swap(order[i], order[(int)((rng >> 33) % (unsigned)(i + 1))]);
                        }
                    }
                    for (int oi = 0; oi < nthr; ++oi) {
                        const int i = order[oi];
                        Fiber *f = g_pool[i];
                        if (f->done) continue;
                        g_cur = f;
                        g_threadIdx = f->tid;
                        swapcontext(&g_main, &f->ctx);
                        if (!f->done) ++remaining;
                    }
                    g_cur = nullptr;
                    if (remaining > 0 && g_progress == before) {
                        g_threadIdx = uint3{0, 0, 0};
                        die("deadlock: every live thread of the CTA is waiting");
                    }
                }
            }
    g_body = nullptr;
}

// ---- fake device memory and driver ---------------------------------------------------------------
void *dev_alloc(size_t bytes)
{
    const size_t n = (bytes + 255) / 256 * 256;
    void *p = aligned_alloc(256, n ? n : 256);
    if (p) memset(p, 0xff, n ? n : 256);   // NaN: device memory is uninitialised
    return p;
}
void dev_free(void *p) { free(p); }

int device_count()
{
    const char *e = getenv("PBX_EMU_NO_DEVICE");
    return (e && e[0] == '1') ? 0 : 1;
}

namespace {
long long g_maps = 0, g_maps3_swz = 0;
CUresult encode_tiled(CUtensorMap *m, CUtensorMapDataType dt, cuuint32_t rank, void *base,
                      const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box,
                      const cuuint32_t *estr, CUtensorMapInterleave, CUtensorMapSwizzle sw,
                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill)
{
    // the constraints of the real encoder that matter for this library
    if (dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT64 || rank < 1 || rank > 3) return CUDA_ERROR_INVALID_VALUE;
    if ((uintptr_t)base & 15) return CUDA_ERROR_INVALID_VALUE;
    ++g_maps;
    if (rank == 3 && sw == CU_TENSOR_MAP_SWIZZLE_128B) ++g_maps3_swz;
    memset((void *)m, 0, sizeof *m);
    m->base = (double *)base;
    m->rank = (int)rank;
    m->swizzle = (int)sw;
    m->stride_bytes[0] = 8;
    for (cuuint32_t d = 0; d < rank; ++d) {
        if (dims[d] == 0 || box[d] == 0 || box[d] > 256 || estr[d] != 1) return CUDA_ERROR_INVALID_VALUE;
        m->dims[d] = dims[d];
        m->box[d] = box[d];
        if (d > 0) {
            if (strides[d - 1] % 16 || strides[d - 1] >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;
            m->stride_bytes[d] = strides[d - 1];
        }
    }
    if ((box[0] * 8) % 16) return CUDA_ERROR_INVALID_VALUE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * 8 > 128) return CUDA_ERROR_INVALID_VALUE;
    return CUDA_SUCCESS;
}
}  // namespace

long long tensor_maps_total() { return g_maps; }
long long tensor_maps3_swizzled_total() { return g_maps3_swz; }

void *driver_entry_point(const char *name)
{
    const char *e = getenv("PBX_EMU_NO_TMA");
    if (e && e[0] == '1') return nullptr;
    if (strcmp(name, "cuTensorMapEncodeTiled") == 0) return (void *)&encode_tiled;
    return nullptr;
}

}  // namespace pbx_emu

extern "C" long long pbx_emu_launches_total(void) { return pbx_emu::launches_total(); }
extern "C" long long pbx_emu_tensor_maps_total(void) { return pbx_emu::tensor_maps_total(); }
extern "C" long long pbx_emu_tensor_maps3_swizzled_total(void) { return pbx_emu::tensor_maps3_swizzled_total(); }
