// Functional model of poissbox_b200/csrc/pbx_ptx.cuh for the CPU kernel-logic harness: same
// names and signatures, implemented on the fiber scheduler of pbx_emu.h.  tests/emu/gen.py
// installs this file as "pbx_ptx.cuh" next to the transformed sources.  TEST INFRASTRUCTURE ONLY.
#pragma once

#include <cuda.h>

#include <sched.h>

#include <cstdint>
#include <ctime>

#include "pbx_emu.h"

namespace pbx {
namespace ptx {

// shared-window address: offset from the CTA's dynamic shared memory base (1 KiB aligned)
inline uint32_t smem_u32(const void *p)
{
    return (uint32_t)((const unsigned char *)p - pbx_emu::dyn_smem());
}

template <int NTHREADS>
inline void named_bar_sync(int id)
{
    pbx_emu::barrier(id, NTHREADS);
}
template <int ID, int NTHREADS>
inline void named_bar_sync_const()
{
    pbx_emu::barrier(ID, NTHREADS);
}

// mbarrier: phase bit, arrival count of the phase, pending arrivals, pending transaction bytes
struct MBar {
    int32_t tx;
    uint16_t pending;
    uint16_t init_phase;   // bit 15: phase, bits 0-14: arrival count per phase
};
static_assert(sizeof(MBar) == 8, "an mbarrier is 64 bits");
inline MBar *mb(uint64_t *bar) { return reinterpret_cast<MBar *>(bar); }
inline void mbar_check_complete(MBar *m)
{
    if (m->pending == 0 && m->tx == 0) {
        m->init_phase ^= 0x8000u;
        m->pending = m->init_phase & 0x7fffu;
        pbx_emu::note_progress();
    }
}
inline void mbar_init(uint64_t *bar, int count)
{
    MBar *m = mb(bar);
    m->tx = 0;
    m->pending = (uint16_t)count;
    m->init_phase = (uint16_t)count;
}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    MBar *m = mb(bar);
    if (m->pending == 0) pbx_emu::die("mbarrier: more arrivals than the phase expects");
    m->tx += (int32_t)bytes;
    m->pending -= 1;
    mbar_check_complete(m);
}
inline void mbar_arrive(uint64_t *bar)
{
    MBar *m = mb(bar);
    if (m->pending == 0) pbx_emu::die("mbarrier: more arrivals than the phase expects");
    m->pending -= 1;
    mbar_check_complete(m);
}
inline void mbar_complete_tx(uint64_t *bar, uint32_t bytes)
{
    MBar *m = mb(bar);
    m->tx -= (int32_t)bytes;
    mbar_check_complete(m);
}
inline bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    return ((mb(bar)->init_phase >> 15) & 1u) != (parity & 1u);
}
inline void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) pbx_emu::yield();
}

// TMA tile copies, performed at issue.  Shared-memory image: dense box, dimension 0 fastest; with
// the 128-byte swizzle the 16-byte piece j of 128-byte row q sits at piece j ^ (q & 7).
inline void tma_copy(void *smem, const CUtensorMap *map, const int *c, bool load)
{
    if (smem_u32(smem) & 127u) pbx_emu::die("TMA shared-memory address not 128-byte aligned");
    const int rank = map->rank;
    const uint32_t b0 = map->box[0], b1 = rank > 1 ? map->box[1] : 1, b2 = rank > 2 ? map->box[2] : 1;
    if (map->swizzle == CU_TENSOR_MAP_SWIZZLE_128B && (b0 * 8 != 128 || (smem_u32(smem) & 1023u)))
        pbx_emu::die("swizzled TMA box: inner extent must be 128 bytes, address 1 KiB aligned");
    double *s = (double *)smem;
    for (uint32_t k = 0; k < b2; ++k)
        for (uint32_t j = 0; j < b1; ++j)
            for (uint32_t i = 0; i < b0; ++i) {
                const long long g0 = c[0] + (long long)i, g1 = rank > 1 ? c[1] + (long long)j : 0,
                                g2 = rank > 2 ? c[2] + (long long)k : 0;
                const bool in = g0 >= 0 && g0 < (long long)map->dims[0] &&
                                (rank < 2 || (g1 >= 0 && g1 < (long long)map->dims[1])) &&
                                (rank < 3 || (g2 >= 0 && g2 < (long long)map->dims[2]));
                size_t off = ((size_t)k * b1 + j) * b0 + i;
                if (map->swizzle == CU_TENSOR_MAP_SWIZZLE_128B) {
                    const size_t row = off / 16, piece = (off % 16) / 2, w = off & 1;
                    off = row * 16 + ((piece ^ (row & 7)) << 1) + w;
                }
                double *g = (double *)((char *)map->base + g0 * 8 +
                                       (rank > 1 ? g1 * (long long)map->stride_bytes[1] : 0) +
                                       (rank > 2 ? g2 * (long long)map->stride_bytes[2] : 0));
                if (load)
                    s[off] = in ? *g : 0.0;
                else if (in)
                    *g = s[off];
            }
}
inline uint32_t tma_box_bytes(const CUtensorMap *map)
{
    uint32_t n = 8;
    for (int d = 0; d < map->rank; ++d) n *= map->box[d];
    return n;
}
inline void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    const int c[3] = {c0, c1, c2};
    if (map->rank != 3) pbx_emu::die("tma_load_3d on a map of another rank");
    tma_copy(dst, map, c, true);
    mbar_complete_tx(bar, tma_box_bytes(map));
    pbx_emu::note_progress();
}
inline void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    const int c[3] = {c0, c1, 0};
    if (map->rank != 2) pbx_emu::die("tma_load_2d on a map of another rank");
    tma_copy(dst, map, c, true);
    mbar_complete_tx(bar, tma_box_bytes(map));
    pbx_emu::note_progress();
}
inline void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1)
{
    const int c[3] = {c0, c1, 0};
    if (map->rank != 2) pbx_emu::die("tma_store_2d on a map of another rank");
    tma_copy(const_cast<void *>(src), map, c, false);
}
inline void tma_commit() {}
inline void tma_wait_read0() {}
inline void tma_wait_all0() {}
inline void fence_proxy_async() {}
inline void fence_mbar_init() {}


inline void prefetch_l2(const void *) {}

// system-scope flags: the "peers" of the harness are other PROCESSES sharing the memory (mmap),
// so these are real atomics; a spinning thread yields the CPU and gives up after two minutes
inline void st_release_sys(unsigned long long *p, unsigned long long v)
{
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
inline unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
inline void st_relaxed_sys(double *p, double v) { *(volatile double *)p = v; }
inline double ld_relaxed_sys(const double *p) { return *(const volatile double *)p; }
inline void fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long spin_start() { return (long long)time(nullptr); }
inline void spin_pause(long long t0)
{
    sched_yield();
    if ((long long)time(nullptr) - t0 > 120) pbx_emu::die("peer flag never arrived");
}

}  // namespace ptx
}  // namespace pbx
