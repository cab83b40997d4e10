// cuda_runtime.h of the CPU kernel-logic harness (tests/emu): the handful of CUDA runtime entry
// points the library uses, modelled on host memory.  TEST INFRASTRUCTURE ONLY -- see pbx_emu.h.
// All functions have C++ linkage and internal names, so nothing here can shadow the real
// libcudart in a process that also loads the product library.
#pragma once

#include <chrono>

#include "pbx_emu.h"

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorNotSupported = 801 };
typedef struct pbx_emu_stream *cudaStream_t;
struct pbx_emu_event {
    std::chrono::steady_clock::time_point t;
};
typedef pbx_emu_event *cudaEvent_t;
enum cudaMemcpyKind {
    cudaMemcpyHostToHost = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3
};
typedef int cudaFuncAttribute;
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaEventDisableTiming = 2 };
enum { cudaEnableDefault = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t {
    char reserved[64];
};

namespace pbx_emu {
void *dev_alloc(size_t bytes);
void dev_free(void *p);
int device_count();
void *driver_entry_point(const char *name);
}  // namespace pbx_emu

template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes)
{
    *p = (T *)pbx_emu::dev_alloc(bytes);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMallocAsync(T **p, size_t bytes, cudaStream_t)
{
    return cudaMalloc(p, bytes);
}
// memory pools: one opaque pool, allocations as cudaMalloc
typedef struct pbx_emu_pool *cudaMemPool_t;
enum cudaMemAllocationType { cudaMemAllocationTypePinned = 1 };
enum cudaMemAllocationHandleType { cudaMemHandleTypeNone = 0 };
enum cudaMemLocationType { cudaMemLocationTypeDevice = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
struct cudaMemLocation {
    cudaMemLocationType type;
    int id;
};
struct cudaMemPoolProps {
    cudaMemAllocationType allocType;
    cudaMemAllocationHandleType handleTypes;
    cudaMemLocation location;
};
static inline cudaError_t cudaMemPoolCreate(cudaMemPool_t *pool, const cudaMemPoolProps *)
{
    *pool = (cudaMemPool_t)(uintptr_t)1;
    return cudaSuccess;
}
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMallocFromPoolAsync(T **p, size_t bytes, cudaMemPool_t, cudaStream_t)
{
    return cudaMalloc(p, bytes);
}
template <class T>
static inline cudaError_t cudaMallocHost(T **p, size_t bytes)
{
    return cudaMalloc(p, bytes);
}
enum { cudaHostAllocMapped = 2 };
template <class T>
static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned)
{
    return cudaMalloc(p, bytes);
}
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned)
{
    *d = h;   // one address space on the harness
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p)
{
    pbx_emu::dev_free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t)
{
    return cudaMemcpy(d, s, n, k);
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n)
{
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { return cudaMemset(d, v, n); }
static inline cudaError_t cudaGetDeviceCount(int *n)
{
    *n = pbx_emu::device_count();
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d)
{
    *d = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "pbx_emu error"; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    *s = nullptr;   // everything runs in program order on the harness
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = new pbx_emu_event();
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete e;
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int)
{
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int)
{
    *v = attr == cudaDevAttrMultiProcessorCount ? 4 : 0;   // a small "GPU": persistent kernels loop
    return cudaSuccess;
}
static inline cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long,
                                                  cudaDriverEntryPointQueryResult *q)
{
    *fn = pbx_emu::driver_entry_point(name);
    if (q) *q = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
