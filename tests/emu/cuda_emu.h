// tests/emu/cuda_emu.h -- a minimal single-OS-thread CUDA *logic* emulator.  TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU, and a gpurun round trip takes minutes, so the kernel sources
// under deltaq_b200/csrc are also compilable with g++ against this shim (-DDQ_EMU) into
// tests/emu/libdeltaq_emu.so.  It is used by tests/ to debug indexing / ranking / scan logic on the CPU.
// It is NOT a product path: deltaq_b200 never loads it, it is not a fallback, and it proves nothing about
// memory ordering or performance.  Parity claims rest on the `-m gpu` tests only.
//
// Model: blocks run one after another in blockIdx order; the threads of a block are fibers switched at
// __syncthreads() and at warp collectives (lockstep only where the code asks for it).  Atomics are plain
// operations.  Because a block runs to completion before the next starts, decoupled look-back never has to
// wait (a spin that would wait forever aborts instead).
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct int2 { int32_t x, y; };
struct __attribute__((aligned(16))) int4 { int32_t x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostRegisterDefault = 0, cudaHostAllocDefault = 0, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; };

namespace emu {

struct Fiber {
    void *sp = nullptr;
    dim3 tid;
    bool done = false;
    char *stack = nullptr;
};
struct WarpSync {
    uint32_t arrived = 0;
    int gen = 0;
    uint64_t val[2][32];
};

constexpr size_t kStackBytes = 256 * 1024;
extern Fiber *cur;
extern void *sched_sp;
extern dim3 g_blockIdx, g_blockDim, g_gridDim;
extern int g_live, g_bar_count, g_bar_gen;
extern std::vector<WarpSync> g_warps;
extern unsigned char *g_dyn_smem;
extern const std::function<void()> *g_body;
extern long g_spin_guard;

extern "C" void emu_switch(void **from_sp, void *to_sp);

inline void yield() { emu_switch(&cur->sp, sched_sp); }
inline unsigned char *dyn_smem() { return g_dyn_smem; }
inline unsigned lane() { return cur->tid.x & 31u; }
inline WarpSync &warp() { return g_warps[cur->tid.x >> 5]; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);

// all lanes named in `mask` deposit a value; returns the snapshot of all 32 slots
inline const uint64_t *exchange(unsigned mask, uint64_t v)
{
    WarpSync &W = warp();
    unsigned l = lane();
    assert((mask >> l) & 1u);
    int g = W.gen;
    uint64_t *buf = W.val[g & 1];
    buf[l] = v;
    W.arrived |= 1u << l;
    if (W.arrived == mask) {
        W.arrived = 0;
        W.gen = g + 1;
    } else {
        while (W.gen == g) yield();
    }
    return buf;
}

inline void spin_wait()
{
    if (++g_spin_guard > 100000000L) {
        fprintf(stderr, "emu: spin wait would never be satisfied (blocks run sequentially)\n");
        abort();
    }
}

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

// ---- synchronisation -----------------------------------------------------------------------------------
static inline void __syncthreads()
{
    int g = emu::g_bar_gen;
    if (++emu::g_bar_count == emu::g_live) {
        emu::g_bar_count = 0;
        emu::g_bar_gen = g + 1;
    } else {
        while (emu::g_bar_gen == g) emu::yield();
    }
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::exchange(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}
static inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }

// ---- warp collectives ----------------------------------------------------------------------------------
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    static_assert(sizeof(T) <= 8, "shfl of <= 8 bytes");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *b = emu::exchange(mask, raw);
    unsigned l = emu::lane();
    unsigned s = (l & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    uint64_t r = ((mask >> s) & 1u) ? b[s] : raw;
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *b = emu::exchange(mask, raw);
    unsigned l = emu::lane();
    unsigned base = l & ~(unsigned)(width - 1);
    uint64_t r = (l - base >= delta && ((mask >> (l - delta)) & 1u)) ? b[l - delta] : raw;
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *b = emu::exchange(mask, raw);
    unsigned l = emu::lane();
    unsigned base = l & ~(unsigned)(width - 1);
    uint64_t r = (l + delta < base + (unsigned)width && ((mask >> (l + delta)) & 1u)) ? b[l + delta] : raw;
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *b = emu::exchange(mask, raw);
    unsigned s = emu::lane() ^ (unsigned)lanemask;
    (void)width;
    uint64_t r = ((mask >> s) & 1u) ? b[s] : raw;
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    const uint64_t *b = emu::exchange(mask, pred ? 1 : 0);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && b[i]) r |= 1u << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *b = emu::exchange(mask, raw);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && b[i] == raw) r |= 1u << i;
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
    const uint64_t *b = emu::exchange(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r += (unsigned)b[i];
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    const uint64_t *b = emu::exchange(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r = std::max(r, (unsigned)b[i]);
    return r;
}

// CUDA's global min/max overloads
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }

// ---- bit intrinsics ------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (shift & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }

// ---- atomics (one OS thread: plain RMW) ------------------------------------------------------------------
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- host runtime ---------------------------------------------------------------------------------------
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// every buffer of the emulator is host memory the "device" can address
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) { a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = const_cast<void *>(p); a->hostPointer = const_cast<void *>(p); return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "deltaq CPU logic emulator");
    p->multiProcessorCount = 4;
    p->totalGlobalMem = (size_t)8 << 30;
    p->major = 10;
    return cudaSuccess;
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
