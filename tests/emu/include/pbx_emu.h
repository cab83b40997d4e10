// pbx_emu.h -- a small functional model of the CUDA execution model on the CPU, so that the
// library's kernels (poissbox_b200/csrc/*.cu, source-transformed by tests/emu/gen.py) can be run
// and checked in a container without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (libpbx.so, the poissbox_b200 package, bench.py)
// includes, links or loads anything under tests/emu; the product has no CPU path and fails with
// PBX_ERR_CUDA when there is no device.  The harness exists to check KERNEL LOGIC (indexing, tile
// geometry, barriers, look-back, boundary handling) on the CPU between the few GPU runs a round
// affords; numerical parity claims are made by the `-m gpu` tests against the real kernels.
//
// Model: one CTA at a time; every CUDA thread of the CTA is a ucontext fiber scheduled round-robin
// and switched at barriers, warp shuffles and mbarrier waits.  Dynamic shared memory is one
// 1 KiB-aligned arena filled with NaN before each CTA; device memory is host memory filled with
// NaN at allocation (reading something never written shows up in the results).  TMA copies are
// performed synchronously at issue.  Not modelled: races (fibers only switch at the points above),
// memory ordering, timing.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace pbx_emu {

extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

// run `body` once per thread of every CTA of the grid
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
unsigned char *dyn_smem();          // base of the CTA's dynamic shared memory (1 KiB aligned)
void yield();                       // let the other threads of the CTA run
void note_progress();               // something observable changed (deadlock detection)
void barrier(int id, int count);    // bar.sync id, count
int linear_tid();
uint64_t shfl_bits(uint64_t v, int src_lane);   // value of lane src_lane (own value if out of range)
long long launches_total();
[[noreturn]] void die(const char *what);

template <class T>
inline T shfl_any(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    b = shfl_bits(b, src);
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace pbx_emu

#define threadIdx (pbx_emu::g_threadIdx)
#define blockIdx (pbx_emu::g_blockIdx)
#define blockDim (pbx_emu::g_blockDim)
#define gridDim (pbx_emu::g_gridDim)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

static inline void __syncthreads()
{
    pbx_emu::barrier(0, (int)(blockDim.x * blockDim.y * blockDim.z));
}
template <class T>
static inline T __ldg(const T *p)
{
    return *p;
}
template <class T>
static inline T __ldcs(const T *p)
{
    return *p;
}
template <class T>
static inline T __ldcg(const T *p)
{
    return *(const volatile T *)p;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32)
{
    return pbx_emu::shfl_any(v, src);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta, int = 32)
{
    const int lane = pbx_emu::linear_tid() & 31;
    return pbx_emu::shfl_any(v, lane + delta > 31 ? lane : lane + delta);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta, int = 32)
{
    const int lane = pbx_emu::linear_tid() & 31;
    return pbx_emu::shfl_any(v, lane - delta < 0 ? lane : lane - delta);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32)
{
    const int lane = pbx_emu::linear_tid() & 31;
    return pbx_emu::shfl_any(v, lane ^ m);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { pbx_emu::shfl_any(0, 0); }
// round-to-nearest intrinsics: plain IEEE operations (the harness is built with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline long long clock64() { return 0; }
static inline void __threadfence() {}
static inline unsigned atomicAdd(unsigned *p, unsigned v)   // one CTA at a time, fibers switch at barriers only
{
    const unsigned o = *p;
    *p = o + v;
    return o;
}
[[noreturn]] static inline void __trap() { pbx_emu::die("__trap()"); }
using std::fma;
