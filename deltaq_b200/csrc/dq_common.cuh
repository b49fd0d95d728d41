// dq_common.cuh -- shared device/host helpers for libdeltaq_cuda (sm_100a).
//
// The same sources also compile with g++ -DDQ_EMU against tests/emu/cuda_emu.h; that build is a
// test-side logic emulator only (see that header) and is never loaded by the product.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

#ifdef DQ_EMU
#include "cuda_emu.h"
#define DQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define DQ_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#define DQ_SPIN_HINT() emu::spin_wait()
#else
#include <cuda_runtime.h>
#define DQ_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define DQ_SPIN_HINT() ((void)0)
#endif

namespace dq {

constexpr int kWarpThreads = 32;
constexpr unsigned kFullMask = 0xffffffffu;

__host__ __device__ __forceinline__ uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }

// ---- accessors for the decoupled look-back descriptors ------------------------------------------------------
// A descriptor is ONE 32/64-bit word holding status and value together, and nothing else is communicated
// through it, so single-word atomicity is all the protocol needs: relaxed gpu-scope accesses (served by L2),
// no fences.  (acquire/release here compiles to CCTL.IVALL + ERRBAR around every poll -- 35 % of the pass
// kernel's stall samples in the first version, profiles/r01_onesweep_pass_full.md.)
#ifdef DQ_EMU
__device__ __forceinline__ uint32_t ld_desc(const uint32_t *p) { return *p; }
__device__ __forceinline__ uint64_t ld_desc(const uint64_t *p) { return *p; }
__device__ __forceinline__ void st_desc(uint32_t *p, uint32_t v) { *p = v; }
__device__ __forceinline__ void st_desc(uint64_t *p, uint64_t v) { *p = v; }
#else
__device__ __forceinline__ uint32_t ld_desc(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_desc(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_desc(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// streaming (evict-first) accessors for data that is touched once per pass
template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
template <typename T> __device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

}  // namespace dq
